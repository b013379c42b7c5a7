/*
 * wrsn_engine.cuh — per-environment simulation engine of the batched WRSN simulator.
 *
 * One CTA advances one environment.  The environment's working set (node rows, event clock,
 * charger records) lives in shared memory for the duration of a launch; thread 0 (the
 * "leader") owns the discrete-event clock, every thread owns the nodes  i = tid, tid+G, ...
 *
 * What is reproduced (reference file:line):
 *   Network.operate / setLevels / check_targets        physical_env/network/Network.py:37-85
 *   Node.operate / send_package / receive_package /
 *   find_receiver / check_status / charger_(dis)connection   physical_env/network/Node.py:45-151
 *   MobileCharger.operate_step / move / move_step /
 *   recharge / charge / charge_step / checkStatus       physical_env/mc/MobileCharger.py:34-140
 *   WRSN.reset / step / update_reward / get_network_fitness / get_reward / translate
 *                                                       rl_env/WRSN.py:41-127,188-227,289-330
 *   and, underneath, SimPy 4.0.1's ordering rule (time, priority, insertion counter) with URGENT
 *   process starts and NORMAL timeouts / completions / conditions, and the nested AnyOf chain of
 *   WRSN.step with its check-callback removal.
 *
 * How it differs from the reference's shape (B200-first, not a translation):
 *   - the N per-node generator processes collapse into ONE block event per half second, executed
 *     node-parallel;
 *   - the per-packet multi-hop recursion collapses into a routing tree (receiver per node) with
 *     per-node relay counts, recomputed only after a death; a node's tick is then the ordered
 *     replay of ITS OWN fp64 operation sequence (relayed packets of lower ids, top-up, own packets,
 *     relayed packets of higher ids), evaluated in closed form when it stays inside one binade
 *     (sub_chain below) — bit-identical to the sequential reference;
 *   - a tick in which some node might die takes an exact serial path (leader thread);
 *   - every charger generator tree is one flat state machine ("process slot") with one pending
 *     event.
 *
 * The same source compiles for the device (nvcc, sm_100a) and, with WRSN_HOST_EMU, as plain
 * single-lane C++ used ONLY by tests/ to check the event logic on a box without a GPU.
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "wrsn_b200.h"

#if defined(WRSN_HOST_EMU)
#define WRSN_HD static inline
#define WRSN_D static inline
#else
#define WRSN_HD __host__ __device__ static inline
#define WRSN_D __device__ static
#endif

/* ------------------------------------------------------------------ layouts */
struct WrsnLayout {
    int64_t off[WRSN_F_COUNT];
    int64_t resident, total;                        /* bytes mirrored in shared memory / bytes per record */
    int64_t s_own, s_scr0, s_scr1, s_bcast, s_red, smem_total;
    int64_t soff[WRSN_S_COUNT];
    int64_t scen_total;
    int32_t scr_len;                                /* doubles per scratch row */
};

WRSN_HD int64_t wrsn_a16(int64_t x) { return (x + 15) & ~(int64_t)15; }

WRSN_HD void wrsn_make_layout(const wrsn_dims *d, WrsnLayout *L) {
    const int64_t Np = d->Npad;
    int64_t o = 0;
    L->off[WRSN_F_HDR] = o; o += wrsn_a16(8 * WRSN_H_LEN);
    L->off[WRSN_F_MC] = o; o += wrsn_a16(8 * (int64_t)(d->M > 0 ? d->M : 1) * WRSN_MC_LEN);
    L->off[WRSN_F_PROC] = o; o += wrsn_a16(8 * (int64_t)d->n_slot * WRSN_PR_LEN);
    L->off[WRSN_F_ENERGY] = o; o += 8 * Np;
    L->off[WRSN_F_RR] = o; o += 8 * Np;
    L->off[WRSN_F_CS] = o; o += 8 * Np;
    L->off[WRSN_F_ESEND] = o; o += 8 * Np;
    L->off[WRSN_F_LOGC] = o; o += 8 * Np;
    L->off[WRSN_F_NBEF] = o; o += 2 * Np;
    L->off[WRSN_F_NAFT] = o; o += 2 * Np;
    L->off[WRSN_F_LEVEL] = o; o += 2 * Np;
    L->off[WRSN_F_PARENT] = o; o += 2 * Np;
    L->off[WRSN_F_STATUS] = o; o += Np;
    L->off[WRSN_F_TACT] = o; o += wrsn_a16(4 * (int64_t)d->Tw);
    L->off[WRSN_F_CONN] = o; o += wrsn_a16(4 * (int64_t)(d->M > 0 ? d->M : 1) * d->W);
    L->resident = o;
    L->off[WRSN_F_LOGTICK] = o; o += 8 * Np;
    L->off[WRSN_F_RING] = o; o += 8 * Np * WRSN_RING;
    L->total = o;
    /* shared-memory extras behind the resident image */
    int64_t Tp = ((int64_t)d->T + 15) & ~(int64_t)15;
    L->scr_len = (int32_t)(Np > Tp ? Np : Tp);
    int64_t s = L->resident;
    L->s_own = s; s += 2 * Np;
    L->s_scr0 = s; s += 8 * (int64_t)L->scr_len;
    L->s_scr1 = s; s += 8 * (int64_t)L->scr_len;
    L->s_bcast = s; s += 64;
    L->s_red = s; s += 8 * 32;
    L->smem_total = s;
    /* scenario record */
    o = 0;
    L->soff[WRSN_S_PAR] = o; o += wrsn_a16(8 * WRSN_P_LEN);
    L->soff[WRSN_S_NX] = o; o += 8 * Np;
    L->soff[WRSN_S_NY] = o; o += 8 * Np;
    L->soff[WRSN_S_BS_ESEND] = o; o += 8 * Np;
    L->soff[WRSN_S_NBR_DIST] = o; o += wrsn_a16(8 * (int64_t)d->Emax);
    L->soff[WRSN_S_NBR_ESEND] = o; o += wrsn_a16(8 * (int64_t)d->Emax);
    L->soff[WRSN_S_NBR_PTR] = o; o += wrsn_a16(4 * (Np + 1));
    L->soff[WRSN_S_TGT_PTR] = o; o += wrsn_a16(4 * (Np + 1));
    L->soff[WRSN_S_NBR_IDX] = o; o += wrsn_a16(4 * (int64_t)d->Emax);
    L->soff[WRSN_S_TGT_IDX] = o; o += wrsn_a16(4 * (int64_t)d->TEmax);
    L->soff[WRSN_S_DIRECT] = o; o += Np;
    L->scen_total = wrsn_a16(o);
}

/* ------------------------------------------------------------------ context */
struct Ctx {
    int tid, G;
    int N, T, M, W, Tw, Npad, n_slot, scr_len;
    /* shared-memory image */
    double *hdr, *mc, *proc;
    double *energy, *rr, *cs, *esend, *logc;
    uint16_t *nbef, *naft, *own;
    int16_t *level, *parent;
    uint8_t *status;
    uint32_t *tact, *conn;
    double *scr0, *scr1;
    int *bcast;
    double *red;
    /* global, per environment */
    double *logtick, *ring;
    /* global, per scenario */
    const double *par, *nx, *ny, *bs_esend, *nbr_dist, *nbr_esend;
    const int32_t *nbr_ptr, *tgt_ptr, *nbr_idx, *tgt_idx;
    const uint8_t *direct;
};

enum { WRSN_URGENT = 0, WRSN_NORMAL = 1 };
enum { K_NONE = 0, K_NET, K_UR, K_NODES, K_UNTIL, K_SLOT, K_COND };
enum {   /* program counter of a charger process slot */
    PC_OP_INIT = 1, PC_MOVE_INIT, PC_MS_INIT, PC_MS_FIRE, PC_MS_DONE, PC_MOVE_DEADWAIT, PC_MOVE_DONE,
    PC_RC_INIT, PC_RC_FIRE, PC_RC_DONE, PC_CH_INIT, PC_CS_INIT, PC_CS_FIRE, PC_CS_DONE, PC_CH_DEADWAIT,
    PC_CH_DONE, PC_OP_DONE
};

WRSN_D void ctx_bind(Ctx &c, const wrsn_dims &d, const WrsnLayout &L, const char *scen_row, char *state_row,
                     char *smem, int tid, int G) {
    c.tid = tid; c.G = G;
    c.N = d.N; c.T = d.T; c.M = d.M; c.W = d.W; c.Tw = d.Tw; c.Npad = d.Npad; c.n_slot = d.n_slot;
    c.scr_len = L.scr_len;
    c.hdr = (double *)(smem + L.off[WRSN_F_HDR]);
    c.mc = (double *)(smem + L.off[WRSN_F_MC]);
    c.proc = (double *)(smem + L.off[WRSN_F_PROC]);
    c.energy = (double *)(smem + L.off[WRSN_F_ENERGY]);
    c.rr = (double *)(smem + L.off[WRSN_F_RR]);
    c.cs = (double *)(smem + L.off[WRSN_F_CS]);
    c.esend = (double *)(smem + L.off[WRSN_F_ESEND]);
    c.logc = (double *)(smem + L.off[WRSN_F_LOGC]);
    c.nbef = (uint16_t *)(smem + L.off[WRSN_F_NBEF]);
    c.naft = (uint16_t *)(smem + L.off[WRSN_F_NAFT]);
    c.level = (int16_t *)(smem + L.off[WRSN_F_LEVEL]);
    c.parent = (int16_t *)(smem + L.off[WRSN_F_PARENT]);
    c.status = (uint8_t *)(smem + L.off[WRSN_F_STATUS]);
    c.tact = (uint32_t *)(smem + L.off[WRSN_F_TACT]);
    c.conn = (uint32_t *)(smem + L.off[WRSN_F_CONN]);
    c.own = (uint16_t *)(smem + L.s_own);
    c.scr0 = (double *)(smem + L.s_scr0);
    c.scr1 = (double *)(smem + L.s_scr1);
    c.bcast = (int *)(smem + L.s_bcast);
    c.red = (double *)(smem + L.s_red);
    c.logtick = (double *)(state_row + L.off[WRSN_F_LOGTICK]);
    c.ring = (double *)(state_row + L.off[WRSN_F_RING]);
    c.par = (const double *)(scen_row + L.soff[WRSN_S_PAR]);
    c.nx = (const double *)(scen_row + L.soff[WRSN_S_NX]);
    c.ny = (const double *)(scen_row + L.soff[WRSN_S_NY]);
    c.bs_esend = (const double *)(scen_row + L.soff[WRSN_S_BS_ESEND]);
    c.nbr_dist = (const double *)(scen_row + L.soff[WRSN_S_NBR_DIST]);
    c.nbr_esend = (const double *)(scen_row + L.soff[WRSN_S_NBR_ESEND]);
    c.nbr_ptr = (const int32_t *)(scen_row + L.soff[WRSN_S_NBR_PTR]);
    c.tgt_ptr = (const int32_t *)(scen_row + L.soff[WRSN_S_TGT_PTR]);
    c.nbr_idx = (const int32_t *)(scen_row + L.soff[WRSN_S_NBR_IDX]);
    c.tgt_idx = (const int32_t *)(scen_row + L.soff[WRSN_S_TGT_IDX]);
    c.direct = (const uint8_t *)(scen_row + L.soff[WRSN_S_DIRECT]);
}

/* ------------------------------------------------------------------ group primitives */
WRSN_D void gsync(const Ctx &c) {
#if !defined(WRSN_HOST_EMU)
    if (c.G == 32) __syncwarp(); else __syncthreads();
#else
    (void)c;
#endif
}

#if !defined(WRSN_HOST_EMU)
#define WRSN_WARP_RED(v, OP)                                                  \
    for (int o_ = 16; o_ > 0; o_ >>= 1) { auto w_ = __shfl_xor_sync(0xffffffffu, v, o_); v = OP(v, w_); }
#define WRSN_OP_ADD(a, b) ((a) + (b))
#define WRSN_OP_OR(a, b) ((a) | (b))
#endif

WRSN_D double red_sum(const Ctx &c, double v) {
#if !defined(WRSN_HOST_EMU)
    WRSN_WARP_RED(v, WRSN_OP_ADD)
    if (c.G == 32) return v;
    __syncthreads();
    if ((c.tid & 31) == 0) c.red[c.tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int k = 0; k < (c.G >> 5); k++) s += c.red[k];
    return s;
#else
    (void)c; return v;
#endif
}
WRSN_D double red_min(const Ctx &c, double v) {
#if !defined(WRSN_HOST_EMU)
    WRSN_WARP_RED(v, fmin)
    if (c.G == 32) return v;
    __syncthreads();
    if ((c.tid & 31) == 0) c.red[c.tid >> 5] = v;
    __syncthreads();
    double s = c.red[0];
    for (int k = 1; k < (c.G >> 5); k++) s = fmin(s, c.red[k]);
    return s;
#else
    (void)c; return v;
#endif
}
WRSN_D int red_or(const Ctx &c, int v) {
#if !defined(WRSN_HOST_EMU)
    if (c.G == 32) return __any_sync(0xffffffffu, v) ? 1 : 0;
    return __syncthreads_or(v) ? 1 : 0;
#else
    (void)c; return v ? 1 : 0;
#endif
}
WRSN_D void atomic_or_u32(uint32_t *p, uint32_t v) {
#if !defined(WRSN_HOST_EMU)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
WRSN_D void atomic_add_i32(int *p, int v) {
#if !defined(WRSN_HOST_EMU)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
WRSN_D void atomic_max_nonneg(double *p, double v) {      /* v >= 0: order of the bit patterns == order of the values */
#if !defined(WRSN_HOST_EMU)
    atomicMax((long long *)p, __double_as_longlong(v));
#else
    if (v > *p) *p = v;
#endif
}

WRSN_D double euclid2(double ax, double ay, double bx, double by) {
    /* scipy.spatial.distance.euclidean == sqrt(dot(u - v, u - v)) for 2-vectors */
    double dx = ax - bx, dy = ay - by;
    return sqrt(dx * dx + dy * dy);
}

/* ------------------------------------------------------------------ exact replay of a node's subtraction chain
 * n_single times (e -= a) followed by n_pair times (e -= b; e -= a), evaluated sequentially in fp64 by the
 * reference.  Inside one binade [2^k, 2^(k+1)) every representable value is a multiple of u = 2^(k-52), so
 * "e - x" rounds to e - rint(x/u)*u (ties aside): each step removes a fixed integer number of ulps and the
 * whole chain is one exact integer multiply.  If the chain would leave the binade (or x/u is an exact tie)
 * the literal loop runs instead.  Order of the steps does not matter inside the binade. */
WRSN_D double wrsn_pow2_biased(int biased) {       /* 2^(biased - 1023) for 1 <= biased <= 2046 */
    uint64_t b = (uint64_t)biased << 52;
#if !defined(WRSN_HOST_EMU)
    return __longlong_as_double((long long)b);
#else
    double r; memcpy(&r, &b, 8); return r;
#endif
}
WRSN_D int wrsn_biased_exp(double e) {
#if !defined(WRSN_HOST_EMU)
    return (int)((((unsigned long long)__double_as_longlong(e)) >> 52) & 0x7ffull);
#else
    uint64_t b; memcpy(&b, &e, 8); return (int)((b >> 52) & 0x7ffull);
#endif
}
WRSN_D double sub_chain(double e, double a, int n_single, double b, int n_pair) {
    int n_a = n_single + n_pair;
    if (n_a == 0) return e;
    int ex = wrsn_biased_exp(e);                     /* e in [2^(ex-1023), 2^(ex-1022)), u = 2^(ex-1075) */
    if (e > 0.0 && ex > 60 && ex < 1900) {
        double lo = wrsn_pow2_biased(ex);
        double inv_u = wrsn_pow2_biased(2098 - ex), u = wrsn_pow2_biased(ex - 52);
        double qa = a * inv_u, qb = b * inv_u;       /* exact scalings */
        double ra = rint(qa), rb = rint(qb);
        bool tie = (fabs(qa - ra) == 0.5) || (n_pair > 0 && fabs(qb - rb) == 0.5);
        double total = ra * (double)n_a + rb * (double)n_pair;
        if (!tie && total < 4503599627370496.0) {    /* < 2^52 ulps: products and sum are exact integers */
            double r = e - total * u;
            if (r >= lo) return r;
        }
    }
    for (int k = 0; k < n_single; k++) e -= a;
    for (int k = 0; k < n_pair; k++) { e -= b; e -= a; }
    return e;
}

/* ------------------------------------------------------------------ event clock helpers (leader only) */
WRSN_D double take_seq(Ctx &c) { double s = c.hdr[WRSN_H_SEQ]; c.hdr[WRSN_H_SEQ] = s + 1.0; return s; }
WRSN_D double *slot_of(Ctx &c, int s) { return c.proc + (size_t)s * WRSN_PR_LEN; }
WRSN_D double *mc_of(Ctx &c, int a) { return c.mc + (size_t)a * WRSN_MC_LEN; }

WRSN_D void slot_sched(Ctx &c, double *p, int pc, int prio, double delay) {
    p[WRSN_PR_PC] = pc; p[WRSN_PR_PRIO] = prio; p[WRSN_PR_T] = c.hdr[WRSN_H_NOW] + delay;
    p[WRSN_PR_SEQ] = take_seq(c); p[WRSN_PR_PENDING] = 1.0;
}

WRSN_D bool ev_before(double t, double p, double s, double bt, double bp, double bs) {
    if (t != bt) return t < bt;
    if (p != bp) return p < bp;
    return s < bs;
}

/* pick the next event: smallest (time, priority, insertion counter) */
WRSN_D void pick_next(Ctx &c, int *kind, int *idx) {
    double *h = c.hdr;
    int bk = K_NONE, bi = 0;
    double bt = 0, bp = 0, bs = 0;
#define WRSN_CAND(K, I, T, P, S)                                                   \
    { double t_ = (T), p_ = (P), s_ = (S);                                         \
      if (bk == K_NONE || ev_before(t_, p_, s_, bt, bp, bs)) { bk = (K); bi = (I); bt = t_; bp = p_; bs = s_; } }
    if (h[WRSN_H_NET_ON] != 0.0) WRSN_CAND(K_NET, 0, h[WRSN_H_NET_T], WRSN_NORMAL, h[WRSN_H_NET_SEQ])
    if (h[WRSN_H_UR_ON] != 0.0) WRSN_CAND(K_UR, 0, h[WRSN_H_UR_T], WRSN_NORMAL, h[WRSN_H_UR_SEQ])
    WRSN_CAND(K_NODES, 0, h[WRSN_H_NODES_T], WRSN_NORMAL, h[WRSN_H_NODES_SEQ])
    if (h[WRSN_H_UNTIL_ON] != 0.0) WRSN_CAND(K_UNTIL, 0, h[WRSN_H_UNTIL_T], WRSN_URGENT, h[WRSN_H_UNTIL_SEQ])
    for (int s = 0; s < c.n_slot; s++) {
        double *p = slot_of(c, s);
        if (p[WRSN_PR_PENDING] != 0.0) WRSN_CAND(K_SLOT, s, p[WRSN_PR_T], p[WRSN_PR_PRIO], p[WRSN_PR_SEQ])
    }
    int nch = (int)h[WRSN_H_CHAIN_N];
    for (int j = 0; j < nch; j++)
        if (h[WRSN_H_COND_PEND + j] != 0.0) WRSN_CAND(K_COND, j, h[WRSN_H_COND_T + j], WRSN_NORMAL, h[WRSN_H_COND_SEQ + j])
#undef WRSN_CAND
    *kind = bk; *idx = bi;
    if (bk != K_NONE) h[WRSN_H_NOW] = bt;
}

/* ------------------------------------------------------------------ Node.log ring: leave the "all ten entries equal
 * logc" shortcut (entries were not written while it held) */
WRSN_D void leave_uniform(Ctx &c) {
    bool uni = c.hdr[WRSN_H_LOG_UNIFORM] >= 10.0 && c.hdr[WRSN_H_LOG_LEN] >= 10.0;
    if (uni)
        for (int i = c.tid; i < c.N; i += c.G) {
            double v = c.logc[i];
            for (int k = 0; k < WRSN_RING; k++) c.ring[(size_t)k * c.Npad + i] = v;
        }
    gsync(c);
    if (c.tid == 0) c.hdr[WRSN_H_LOG_UNIFORM] = 0.0;
    gsync(c);
}

/* ------------------------------------------------------------------ Network.setLevels + check_targets (Network.py:37-66,84)
 * plus the routing tree the drain tick replays: receiver (Node.find_receiver :92-100), e_send, relay counts,
 * and the per-tick log_energy of every node. */
WRSN_D void do_bfs(Ctx &c) {
    leave_uniform(c);
    const int N = c.N;
    int *cnt = (int *)c.scr0;                       /* 2 ints per node: relayed packets from lower / higher ids */
    for (int i = c.tid; i < N; i += c.G) {
        c.level[i] = (c.status[i] == 1 && c.direct[i]) ? 1 : -1;
        cnt[2 * i] = 0; cnt[2 * i + 1] = 0;
    }
    for (int w = c.tid; w < c.Tw; w += c.G) c.tact[w] = 0u;
    gsync(c);
    for (int cur = 1;; cur++) {
        int any = 0;
        for (int i = c.tid; i < N; i += c.G) {
            if (c.level[i] != cur) continue;
            for (int e = c.tgt_ptr[i]; e < c.tgt_ptr[i + 1]; e++) {
                int t = c.tgt_idx[e];
                atomic_or_u32(&c.tact[t >> 5], 1u << (t & 31));
            }
            for (int e = c.nbr_ptr[i]; e < c.nbr_ptr[i + 1]; e++) {
                int j = c.nbr_idx[e];
                if (c.status[j] == 1 && c.level[j] == -1) { c.level[j] = (int16_t)(cur + 1); any = 1; }
            }
        }
        gsync(c);
        if (!red_or(c, any)) break;
    }
    /* alive = min(targets_active) */
    int dead_t = 0;
    for (int w = c.tid; w < c.Tw; w += c.G) {
        int bits = c.T - 32 * w; if (bits > 32) bits = 32;
        uint32_t full = bits == 32 ? 0xffffffffu : ((1u << bits) - 1u);
        if ((c.tact[w] & full) != full) dead_t = 1;
    }
    dead_t = red_or(c, dead_t);
    /* receivers */
    for (int i = c.tid; i < N; i += c.G) {
        int par = -1; double es = 0.0;
        if (c.status[i] == 1 && c.level[i] >= 1) {
            if (c.direct[i]) { par = -2; es = c.bs_esend[i]; }
            else {
                double bd = 0.0; int lv = c.level[i];
                for (int e = c.nbr_ptr[i]; e < c.nbr_ptr[i + 1]; e++) {
                    int j = c.nbr_idx[e];
                    if (c.level[j] < lv && c.status[j] == 1) {
                        double dd = c.nbr_dist[e];
                        if (par < 0 || dd < bd) { par = j; bd = dd; es = c.nbr_esend[e]; }   /* np.argmin: first minimum */
                    }
                }
            }
        }
        c.parent[i] = (int16_t)par; c.esend[i] = es;
    }
    gsync(c);
    /* relay counts: every packet of source s crosses all its ancestors */
    for (int s = c.tid; s < N; s += c.G) {
        int ow = c.own[s];
        if (c.status[s] != 1 || ow == 0 || c.parent[s] == -1) continue;
        for (int h = c.parent[s]; h >= 0; h = c.parent[h]) atomic_add_i32(&cnt[2 * h + (s < h ? 0 : 1)], ow);
    }
    gsync(c);
    const double er = c.par[WRSN_P_ERECV];
    for (int i = c.tid; i < N; i += c.G) {
        int nb = cnt[2 * i], na = cnt[2 * i + 1];
        c.nbef[i] = (uint16_t)nb; c.naft[i] = (uint16_t)na;
        double lg = 0.0, es = c.esend[i];
        if (c.status[i] == 1 && c.parent[i] != -1) {
            for (int k = 0; k < nb; k++) { lg += es; lg += er; }
            int ow = c.own[i];
            for (int k = 0; k < ow; k++) lg += es;
            for (int k = 0; k < na; k++) { lg += es; lg += er; }
        }
        c.logc[i] = lg;
    }
    gsync(c);
    if (c.tid == 0) {
        c.hdr[WRSN_H_ALIVE] = dead_t ? 0.0 : 1.0;
        c.hdr[WRSN_H_BFS_DIRTY] = 0.0;
        c.hdr[WRSN_H_NBFS] += 1.0;
    }
    gsync(c);
}

/* ------------------------------------------------------------------ Node.operate, k+0.5 tick (Node.py:57-62,92-132) */
WRSN_D void check_status_node(Ctx &c, int i) {     /* Node.py:148-151 */
    if (c.energy[i] <= c.par[WRSN_P_THR]) { c.status[i] = 0; c.cs[i] = 0.0; }
}

/* exact serial tick: packet by packet, hop by hop, as the reference does it (leader only) */
WRSN_D int drain_serial(Ctx &c) {
    const int N = c.N;
    const double thr = c.par[WRSN_P_THR], cap = c.par[WRSN_P_CAP], er = c.par[WRSN_P_ERECV];
    int deaths = 0;
    for (int i = 0; i < N; i++) c.logtick[i] = 0.0;
    for (int i = 0; i < N; i++) {
        if (c.status[i] == 0) continue;
        c.energy[i] = fmin(c.energy[i] + c.rr[i] * 0.5, cap);
        int ow = c.own[i];
        for (int k = 0; k < ow; k++) {
            int h = i; bool pay_recv = false;
            for (;;) {
                if (pay_recv) {                      /* receive_package */
                    if (c.energy[h] - thr < er) {
                        c.energy[h] = thr;
                        if (c.status[h] == 1) deaths++;
                        check_status_node(c, h);
                        break;
                    }
                    c.energy[h] -= er;
                }
                int recv = -1; double es = 0.0;      /* send_package */
                if (c.direct[h]) { recv = -2; es = c.bs_esend[h]; }
                else {
                    double bd = 0.0; int lv = c.level[h];
                    for (int e = c.nbr_ptr[h]; e < c.nbr_ptr[h + 1]; e++) {
                        int j = c.nbr_idx[e];
                        if (c.level[j] < lv && c.status[j] == 1) {
                            double dd = c.nbr_dist[e];
                            if (recv < 0 || dd < bd) { recv = j; bd = dd; es = c.nbr_esend[e]; }
                        }
                    }
                }
                bool sent = false;
                if (recv != -1) {
                    if (c.energy[h] - thr < es) c.energy[h] = thr;
                    else { c.energy[h] -= es; sent = true; }
                }
                if (sent) c.logtick[h] += es;
                if (pay_recv) c.logtick[h] += er;
                if (c.status[h] == 1 && c.energy[h] <= thr) deaths++;
                check_status_node(c, h);
                if (!sent || recv == -2) break;
                h = recv; pay_recv = true;
            }
        }
    }
    return deaths;
}

WRSN_D void ev_nodes_drain(Ctx &c) {
    const int N = c.N;
    const double thr = c.par[WRSN_P_THR], cap = c.par[WRSN_P_CAP], er = c.par[WRSN_P_ERECV];
    const double slack = 1e-6;
    int slow = c.hdr[WRSN_H_BFS_DIRTY] != 0.0 ? 1 : 0;   /* routing tree is stale (Network.operate has stopped): serial path */
    for (int i = c.tid; i < N; i += c.G) {
        if (c.status[i] != 1) continue;
        double e = c.energy[i], es = c.esend[i];
        int nb = c.nbef[i], na = c.naft[i];
        int ow = c.parent[i] != -1 ? (int)c.own[i] : 0;
        double e1 = sub_chain(e, es, 0, er, nb);
        if (nb > 0 && !(e1 - thr >= slack)) slow = 1;
        double e2 = fmin(e1 + c.rr[i] * 0.5, cap);
        double e3 = sub_chain(e2, es, ow, er, na);
        if (ow + na > 0 && !(e3 - thr >= slack)) slow = 1;
        c.scr0[i] = e3;
    }
    gsync(c);
    slow = red_or(c, slow);
    if (!slow) {
        for (int i = c.tid; i < N; i += c.G)
            if (c.status[i] == 1) c.energy[i] = c.scr0[i];
        gsync(c);
        return;
    }
    leave_uniform(c);
    if (c.tid == 0) {
        int deaths = drain_serial(c);
        c.hdr[WRSN_H_LOG_LITERAL] = 1.0;
        c.hdr[WRSN_H_NSLOW] += 1.0;
        if (deaths > 0) c.hdr[WRSN_H_BFS_DIRTY] = 1.0;
    }
    gsync(c);
}

/* ------------------------------------------------------------------ Node.operate, k+1.0 tick (Node.py:65-77) */
WRSN_D void ev_nodes_book(Ctx &c) {
    const int N = c.N;
    const double cap = c.par[WRSN_P_CAP];
    const int L = (int)c.hdr[WRSN_H_LOG_LEN], head = (int)c.hdr[WRSN_H_LOG_HEAD];
    const bool literal = c.hdr[WRSN_H_LOG_LITERAL] != 0.0;
    const bool uni = !literal && L >= WRSN_RING && c.hdr[WRSN_H_LOG_UNIFORM] >= 10.0;
    for (int i = c.tid; i < N; i += c.G) {
        if (c.status[i] != 1) continue;
        c.energy[i] = fmin(c.energy[i] + c.rr[i] * 0.5, cap);
        double lg = literal ? c.logtick[i] : c.logc[i];
        if (L < WRSN_RING) {
            c.cs[i] = (c.cs[i] * (double)L + lg) / (double)(L + 1);
            c.ring[(size_t)L * c.Npad + i] = lg;
        } else {
            double old = uni ? lg : c.ring[(size_t)head * c.Npad + i];
            c.cs[i] = (c.cs[i] * (double)L - old + lg) / (double)L;
            if (!uni) c.ring[(size_t)head * c.Npad + i] = lg;
        }
    }
    gsync(c);
    if (c.tid == 0) {
        if (L < WRSN_RING) c.hdr[WRSN_H_LOG_LEN] = L + 1;
        else c.hdr[WRSN_H_LOG_HEAD] = (head + 1) % WRSN_RING;
        if (literal) { c.hdr[WRSN_H_LOG_UNIFORM] = 0.0; c.hdr[WRSN_H_LOG_LITERAL] = 0.0; }
        else if (c.hdr[WRSN_H_LOG_UNIFORM] < 1e9) c.hdr[WRSN_H_LOG_UNIFORM] += 1.0;
        c.hdr[WRSN_H_NTICKS] += 1.0;
    }
    gsync(c);
}

/* ------------------------------------------------------------------ WRSN.update_reward (WRSN.py:100-127) */
WRSN_D double charge_rate_to(Ctx &c, const double *m, int node) {   /* alpha / (d + beta) ** 2 */
    double t = euclid2(c.nx[node], c.ny[node], m[WRSN_MC_X], m[WRSN_MC_Y]) + c.par[WRSN_P_MC_BETA];
    return c.par[WRSN_P_MC_ALPHA] / (t * t);
}

WRSN_D void ev_update_reward(Ctx &c) {
    bool any = false;
    for (int a = 0; a < c.M; a++) {
        const double *m = mc_of(c, a);
        if (m[WRSN_MC_STATUS] != 0.0 && m[WRSN_MC_TYPE] != 0.0) any = true;
    }
    if (!any) return;                                /* the priority vector has no other reader */
    const int N = c.N;
    const double thr = c.par[WRSN_P_THR], cap = c.par[WRSN_P_CAP], eps = c.par[WRSN_P_EPSENV];
    double s = 0.0;
    for (int i = c.tid; i < N; i += c.G) {
        double p = c.status[i] != 0 ? c.cs[i] / (c.energy[i] - thr + eps) : 0.0;
        c.scr0[i] = p; s += p;
    }
    double mean = red_sum(c, s) / (double)N;
    s = 0.0;
    for (int i = c.tid; i < N; i += c.G) { double x = c.scr0[i] - mean; s += x * x; }
    double sd = sqrt(red_sum(c, s) / (double)N);
    if (sd == 0.0) sd = eps;
    s = 0.0;
    for (int i = c.tid; i < N; i += c.G) { double q = exp((c.scr0[i] - mean) / sd); c.scr0[i] = q; s += q; }
    double tot = red_sum(c, s);
    if (tot == 0.0) tot = eps;
    gsync(c);
    if (c.tid == 0) {
        for (int a = 0; a < c.M; a++) {
            double *m = mc_of(c, a);
            if (m[WRSN_MC_STATUS] == 0.0 || m[WRSN_MC_TYPE] == 0.0) continue;
            double incentive = 0.0;
            const uint32_t *cm = c.conn + (size_t)a * c.W;
            for (int w = 0; w < c.W; w++) {
                uint32_t bits = cm[w];
                while (bits) {
                    int b = 0; while (!((bits >> b) & 1u)) b++;
                    bits &= bits - 1u;
                    int i = 32 * w + b;
                    if (c.status[i] != 1) continue;
                    double rate = charge_rate_to(c, m, i);
                    double e_no = fmin(c.energy[i] - c.cs[i], thr);
                    double e_with = fmax(c.energy[i] - c.cs[i] + rate, cap);
                    incentive += (c.scr0[i] / tot) * (e_with - e_no) / c.par[WRSN_P_MC_AB2];
                }
            }
            m[WRSN_MC_EXCL] += incentive;
        }
    }
    gsync(c);
}

/* ------------------------------------------------------------------ WRSN.get_network_fitness (WRSN.py:188-220) -> min */
WRSN_D double do_fitness(Ctx &c, double *per_target /* global, may be NULL */) {
    const int N = c.N;
    const double thr = c.par[WRSN_P_THR];
    double *node_t = c.scr0, *lt = c.scr1;
    for (int i = c.tid; i < N; i += c.G) {
        double v = -1.0, l = 0.0;
        if (c.status[i] == 1) {
            l = (c.cs[i] == 0.0) ? INFINITY : (c.energy[i] - thr) / c.cs[i];
            if (c.direct[i]) v = l;
        }
        node_t[i] = v; lt[i] = l;
    }
    gsync(c);
    for (;;) {                                       /* widest path to the base station; only min / max, so any order is exact */
        int changed = 0;
        for (int i = c.tid; i < N; i += c.G) {
            if (c.status[i] != 1 || c.direct[i]) continue;
            double best = -1.0;
            for (int e = c.nbr_ptr[i]; e < c.nbr_ptr[i + 1]; e++) {
                int j = c.nbr_idx[e];
                if (c.status[j] == 1) { double v = node_t[j]; if (v > best) best = v; }
            }
            if (best >= 0.0) {
                double nv = fmin(lt[i], best);
                if (nv > node_t[i]) { node_t[i] = nv; changed = 1; }
            }
        }
        gsync(c);
        if (!red_or(c, changed)) break;
    }
    double *tt = c.scr1;                             /* lt no longer needed */
    gsync(c);
    for (int t = c.tid; t < c.T; t += c.G) tt[t] = 0.0;
    gsync(c);
    for (int i = c.tid; i < N; i += c.G) {
        double v = node_t[i];
        if (v <= 0.0) continue;
        for (int e = c.tgt_ptr[i]; e < c.tgt_ptr[i + 1]; e++) atomic_max_nonneg(&tt[c.tgt_idx[e]], v);
    }
    gsync(c);
    double mn = INFINITY;
    for (int t = c.tid; t < c.T; t += c.G) {
        double v = tt[t];
        if (per_target) per_target[t] = v;
        mn = fmin(mn, v);
    }
    mn = red_min(c, mn);
    gsync(c);
    return mn;
}

/* ------------------------------------------------------------------ chargers (MobileCharger.py) */
WRSN_D void mc_check_status(Ctx &c, double *m) {   /* :134-140 */
    if (m[WRSN_MC_ENERGY] <= c.par[WRSN_P_MC_THR]) { m[WRSN_MC_STATUS] = 0.0; m[WRSN_MC_ENERGY] = c.par[WRSN_P_MC_THR]; }
}

/* all threads: bitmask of nodes with d(node, (x, y)) <= charging_range */
WRSN_D void near_mask(Ctx &c, double x, double y, uint32_t *mask) {
    for (int w = c.tid; w < c.W; w += c.G) mask[w] = 0u;
    gsync(c);
    const double R = c.par[WRSN_P_MC_R];
    for (int i = c.tid; i < c.N; i += c.G)
        if (euclid2(c.nx[i], c.ny[i], x, y) <= R) atomic_or_u32(&mask[i >> 5], 1u << (i & 31));
    gsync(c);
}

#define WRSN_FOR_BITS(mask, W, i)                                              \
    for (int w_ = 0; w_ < (W); w_++)                                           \
        for (uint32_t bits_ = (mask)[w_]; bits_; bits_ &= bits_ - 1u)          \
            for (int i = 32 * w_ + wrsn_ctz(bits_), once_ = 1; once_; once_ = 0)

WRSN_D int wrsn_ctz(uint32_t v) {
#if !defined(WRSN_HOST_EMU)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

/* leader: the move loop head (MobileCharger.move :85-96) */
WRSN_D void mc_move_loop(Ctx &c, double *p, double *m) {
    if (p[WRSN_PR_MT] <= 0.0) { slot_sched(c, p, PC_MOVE_DONE, WRSN_NORMAL, 0.0); return; }
    if (m[WRSN_MC_STATUS] == 0.0) { slot_sched(c, p, PC_MOVE_DEADWAIT, WRSN_NORMAL, p[WRSN_PR_MT]); return; }
    const double v = c.par[WRSN_P_MC_V];
    p[WRSN_PR_MT] = euclid2(p[WRSN_PR_DESTX], p[WRSN_PR_DESTY], m[WRSN_MC_X], m[WRSN_MC_Y]) / v;
    double span = fmin(fmin(p[WRSN_PR_MT], 1.0), (m[WRSN_MC_ENERGY] - c.par[WRSN_P_MC_THR]) / c.par[WRSN_P_MC_PMV]);
    p[WRSN_PR_SPAN] = span;
    p[WRSN_PR_SVX] = p[WRSN_PR_VX] / p[WRSN_PR_TOTAL] * span;
    p[WRSN_PR_SVY] = p[WRSN_PR_VY] / p[WRSN_PR_TOTAL] * span;
    slot_sched(c, p, PC_MS_INIT, WRSN_URGENT, 0.0);
}

/* leader: the charge loop head (MobileCharger.charge :59-72) */
WRSN_D void mc_charge_loop(Ctx &c, double *p, double *m) {
    if (p[WRSN_PR_CHTMP] == 0.0) { slot_sched(c, p, PC_CH_DONE, WRSN_NORMAL, 0.0); return; }
    if (m[WRSN_MC_STATUS] == 0.0) {
        m[WRSN_MC_CPA2] = 0.0;
        slot_sched(c, p, PC_CH_DEADWAIT, WRSN_NORMAL, p[WRSN_PR_CHTMP]);
        return;
    }
    double span = fmin(p[WRSN_PR_CHTMP], 1.0);
    if (m[WRSN_MC_RATE] != 0.0) span = fmin(span, (m[WRSN_MC_ENERGY] - c.par[WRSN_P_MC_THR]) / m[WRSN_MC_RATE]);
    p[WRSN_PR_CHSPAN] = span;
    slot_sched(c, p, PC_CS_INIT, WRSN_URGENT, 0.0);
}

WRSN_D void cond_check(Ctx &c, int j) {            /* simpy Condition._check for AnyOf */
    double *h = c.hdr;
    if (h[WRSN_H_COND_TRIG + j] != 0.0) return;
    h[WRSN_H_COND_TRIG + j] = 1.0;
    h[WRSN_H_COND_PEND + j] = 1.0;
    h[WRSN_H_COND_T + j] = h[WRSN_H_NOW];
    h[WRSN_H_COND_SEQ + j] = take_seq(c);
}

/* one event of a charger process slot.  Called by ALL threads (node-parallel pieces inside). */
WRSN_D void ev_slot(Ctx &c, int s) {
    double *p = slot_of(c, s);
    const int a = (int)p[WRSN_PR_AGENT];
    double *m = mc_of(c, a);
    const int pc = (int)p[WRSN_PR_PC];
    uint32_t *cm = c.conn + (size_t)a * c.W;
    gsync(c);
    if (c.tid == 0) p[WRSN_PR_PENDING] = 0.0;
    switch (pc) {
    case PC_OP_INIT: {                               /* MobileCharger.operate_step :105-132, up to the first yield */
        uint32_t *near = (uint32_t *)c.scr1;
        near_mask(c, p[WRSN_PR_PHY0], p[WRSN_PR_PHY1], near);
        if (c.tid == 0) {
            const double dx = p[WRSN_PR_PHY0], dy = p[WRSN_PR_PHY1], ct = p[WRSN_PR_PHY2];
            const double pm = c.par[WRSN_P_MC_PM], beta = c.par[WRSN_P_MC_BETA], alpha = c.par[WRSN_P_MC_ALPHA];
            double used = euclid2(dx, dy, m[WRSN_MC_X], m[WRSN_MC_Y]) * pm;
            double tmp = 0.0;
            WRSN_FOR_BITS(near, c.W, i) {
                if (c.status[i] == 1) {
                    double t = euclid2(dx, dy, c.nx[i], c.ny[i]) + beta;
                    tmp += alpha / (t * t);
                }
            }
            used += tmp * ct;
            used += euclid2(dx, dy, c.par[WRSN_P_BSX], c.par[WRSN_P_BSY]) * pm;
            m[WRSN_MC_CPA0] = dx; m[WRSN_MC_CPA1] = dy; m[WRSN_MC_CPA2] = ct;
            m[WRSN_MC_TYPE] = 0.0;
            if (used > m[WRSN_MC_ENERGY] - c.par[WRSN_P_MC_THR] - c.par[WRSN_P_MC_CAP200]) {
                p[WRSN_PR_STAGE] = 1.0; p[WRSN_PR_DESTX] = c.par[WRSN_P_BSX]; p[WRSN_PR_DESTY] = c.par[WRSN_P_BSY];
            } else {
                p[WRSN_PR_STAGE] = 3.0; p[WRSN_PR_DESTX] = dx; p[WRSN_PR_DESTY] = dy;
            }
            slot_sched(c, p, PC_MOVE_INIT, WRSN_URGENT, 0.0);
        }
        break;
    }
    case PC_MOVE_INIT:                               /* MobileCharger.move :82-84 */
        if (c.tid == 0) {
            p[WRSN_PR_MT] = euclid2(p[WRSN_PR_DESTX], p[WRSN_PR_DESTY], m[WRSN_MC_X], m[WRSN_MC_Y]) / c.par[WRSN_P_MC_V];
            p[WRSN_PR_VX] = p[WRSN_PR_DESTX] - m[WRSN_MC_X];
            p[WRSN_PR_VY] = p[WRSN_PR_DESTY] - m[WRSN_MC_Y];
            p[WRSN_PR_TOTAL] = p[WRSN_PR_MT];
            mc_move_loop(c, p, m);
        }
        break;
    case PC_MS_INIT:                                 /* move_step :76 */
        if (c.tid == 0) slot_sched(c, p, PC_MS_FIRE, WRSN_NORMAL, p[WRSN_PR_SPAN]);
        break;
    case PC_MS_FIRE:                                 /* move_step :77-78 */
        if (c.tid == 0) {
            m[WRSN_MC_X] = m[WRSN_MC_X] + p[WRSN_PR_SVX];
            m[WRSN_MC_Y] = m[WRSN_MC_Y] + p[WRSN_PR_SVY];
            m[WRSN_MC_ENERGY] -= c.par[WRSN_P_MC_PM] * p[WRSN_PR_SPAN] * c.par[WRSN_P_MC_V];
            slot_sched(c, p, PC_MS_DONE, WRSN_NORMAL, 0.0);
        }
        break;
    case PC_MS_DONE:                                 /* move :95-96 */
        if (c.tid == 0) {
            p[WRSN_PR_MT] -= p[WRSN_PR_SPAN];
            mc_check_status(c, m);
            mc_move_loop(c, p, m);
        }
        break;
    case PC_MOVE_DEADWAIT:
        if (c.tid == 0) slot_sched(c, p, PC_MOVE_DONE, WRSN_NORMAL, 0.0);
        break;
    case PC_MOVE_DONE:                               /* back in operate_step */
        if (c.tid == 0) {
            if (p[WRSN_PR_STAGE] == 1.0) slot_sched(c, p, PC_RC_INIT, WRSN_URGENT, 0.0);
            else {
                m[WRSN_MC_TYPE] = 1.0;
                p[WRSN_PR_CHTMP] = p[WRSN_PR_PHY2];
                slot_sched(c, p, PC_CH_INIT, WRSN_URGENT, 0.0);
            }
        }
        break;
    case PC_RC_INIT:                                 /* recharge :99-103 */
        if (c.tid == 0) {
            if (euclid2(m[WRSN_MC_X], m[WRSN_MC_Y], c.par[WRSN_P_BSX], c.par[WRSN_P_BSY]) <= c.par[WRSN_P_MC_EPS]) {
                m[WRSN_MC_X] = c.par[WRSN_P_BSX]; m[WRSN_MC_Y] = c.par[WRSN_P_BSY];
                m[WRSN_MC_ENERGY] = c.par[WRSN_P_MC_CAP];
            }
            slot_sched(c, p, PC_RC_FIRE, WRSN_NORMAL, 0.0);
        }
        break;
    case PC_RC_FIRE:
        if (c.tid == 0) slot_sched(c, p, PC_RC_DONE, WRSN_NORMAL, 0.0);
        break;
    case PC_RC_DONE:
        if (c.tid == 0) {
            p[WRSN_PR_STAGE] = 3.0; p[WRSN_PR_DESTX] = p[WRSN_PR_PHY0]; p[WRSN_PR_DESTY] = p[WRSN_PR_PHY1];
            slot_sched(c, p, PC_MOVE_INIT, WRSN_URGENT, 0.0);
        }
        break;
    case PC_CH_INIT: {                               /* charge :53-58 */
        near_mask(c, m[WRSN_MC_X], m[WRSN_MC_Y], cm);
        if (c.tid == 0) {
            m[WRSN_MC_CHTIME] = p[WRSN_PR_CHTMP];
            int n = 0;
            for (int w = 0; w < c.W; w++) {
#if !defined(WRSN_HOST_EMU)
                n += __popc(cm[w]);
#else
                n += __builtin_popcount(cm[w]);
#endif
            }
            m[WRSN_MC_NCONN] = n;
            mc_charge_loop(c, p, m);
        }
        break;
    }
    case PC_CS_INIT:                                 /* charge_step :40-44 + Node.charger_connection :134-139 */
        if (c.tid == 0) {
            WRSN_FOR_BITS(cm, c.W, i) {
                if (c.status[i] == 0) continue;
                double r = charge_rate_to(c, m, i);
                c.rr[i] += r; m[WRSN_MC_RATE] += r;
            }
            slot_sched(c, p, PC_CS_FIRE, WRSN_NORMAL, p[WRSN_PR_CHSPAN]);
        }
        break;
    case PC_CS_FIRE:                                 /* charge_step :45-50 + Node.charger_disconnection :141-146 */
        if (c.tid == 0) {
            const double t = p[WRSN_PR_CHSPAN];
            m[WRSN_MC_ENERGY] = m[WRSN_MC_ENERGY] - m[WRSN_MC_RATE] * t;
            m[WRSN_MC_CPA2] = fmax(0.0, m[WRSN_MC_CPA2] - t);
            WRSN_FOR_BITS(cm, c.W, i) {
                if (c.status[i] == 0) continue;
                double r = charge_rate_to(c, m, i);
                c.rr[i] -= r; m[WRSN_MC_RATE] -= r;
            }
            m[WRSN_MC_RATE] = 0.0;
            slot_sched(c, p, PC_CS_DONE, WRSN_NORMAL, 0.0);
        }
        break;
    case PC_CS_DONE:                                 /* charge :69-71 */
        if (c.tid == 0) {
            p[WRSN_PR_CHTMP] -= p[WRSN_PR_CHSPAN];
            m[WRSN_MC_CHTIME] = p[WRSN_PR_CHTMP];
            mc_check_status(c, m);
            mc_charge_loop(c, p, m);
        }
        break;
    case PC_CH_DEADWAIT:
        if (c.tid == 0) slot_sched(c, p, PC_CH_DONE, WRSN_NORMAL, 0.0);
        break;
    case PC_CH_DONE:                                 /* operate_step returns */
        if (c.tid == 0) slot_sched(c, p, PC_OP_DONE, WRSN_NORMAL, 0.0);
        break;
    case PC_OP_DONE:                                 /* the process event itself: callbacks = condition checks */
        if (c.tid == 0) {
            p[WRSN_PR_PROCESSED] = 1.0;
            if (p[WRSN_PR_CURRENT] == 0.0) p[WRSN_PR_USED] = 0.0;     /* superseded process: nobody holds it any more */
            const int nch = (int)c.hdr[WRSN_H_CHAIN_N], det = (int)c.hdr[WRSN_H_CHAIN_DETACH];
            for (int j = 0; j < nch; j++)
                if ((int)c.hdr[WRSN_H_CHAIN_SLOT + j] == s && j > det) cond_check(c, j);
        }
        break;
    default:
        if (c.tid == 0) c.hdr[WRSN_H_ERR] = 2.0;
        break;
    }
    gsync(c);
}

/* a condition event of the AnyOf chain; returns via bcast[2] whether run() stops */
WRSN_D void ev_cond(Ctx &c, int j) {
    if (c.tid == 0) {
        double *h = c.hdr;
        h[WRSN_H_COND_PEND + j] = 0.0;
        const int nch = (int)h[WRSN_H_CHAIN_N];
        /* _build_value: remove the check callbacks of this condition and, recursively, of the nested ones */
        if ((double)j > h[WRSN_H_CHAIN_DETACH]) h[WRSN_H_CHAIN_DETACH] = (double)j;
        if (j + 1 < nch) { if ((double)(j + 1) > h[WRSN_H_CHAIN_DETACH]) cond_check(c, j + 1); }
        else c.bcast[2] = 1;                         /* StopSimulation */
    }
    gsync(c);
}

/* ------------------------------------------------------------------ the event loop: env.run(...) */
WRSN_D void run_loop(Ctx &c) {
    if (c.tid == 0) c.bcast[2] = 0;
    gsync(c);
    for (long guard = 0;; guard++) {
        if (c.tid == 0) {
            int kind, idx;
            pick_next(c, &kind, &idx);
            if (guard > 200000000L) { kind = K_NONE; }
            c.bcast[0] = kind; c.bcast[1] = idx;
            c.hdr[WRSN_H_NEVENTS] += 1.0;
        }
        gsync(c);
        const int kind = c.bcast[0], idx = c.bcast[1];
        const bool net_levels = c.hdr[WRSN_H_NET_STATE] == 1.0, dirty = c.hdr[WRSN_H_BFS_DIRTY] != 0.0;
        const bool drain_phase = c.hdr[WRSN_H_NODES_PHASE] == 1.0;
        gsync(c);
        switch (kind) {
        case K_NET:                                  /* Network.operate :74-80 */
            if (net_levels) {
                if (dirty) do_bfs(c);
                if (c.tid == 0) {
                    c.hdr[WRSN_H_NET_T] = c.hdr[WRSN_H_NOW] + 0.9; c.hdr[WRSN_H_NET_SEQ] = take_seq(c);
                    c.hdr[WRSN_H_NET_STATE] = 2.0;
                }
            } else if (c.tid == 0) {
                if (c.hdr[WRSN_H_ALIVE] == 0.0 || c.hdr[WRSN_H_NOW] >= c.par[WRSN_P_MAXTIME]) c.hdr[WRSN_H_NET_ON] = 0.0;
                else {
                    c.hdr[WRSN_H_NET_T] = c.hdr[WRSN_H_NOW] + 0.1; c.hdr[WRSN_H_NET_SEQ] = take_seq(c);
                    c.hdr[WRSN_H_NET_STATE] = 1.0;
                }
            }
            break;
        case K_UR:
            ev_update_reward(c);
            if (c.tid == 0) { c.hdr[WRSN_H_UR_T] = c.hdr[WRSN_H_NOW] + 1.0; c.hdr[WRSN_H_UR_SEQ] = take_seq(c); }
            break;
        case K_NODES:
            if (drain_phase) ev_nodes_drain(c); else ev_nodes_book(c);
            if (c.tid == 0) {
                c.hdr[WRSN_H_NODES_PHASE] = drain_phase ? 2.0 : 1.0;
                c.hdr[WRSN_H_NODES_T] = c.hdr[WRSN_H_NOW] + 0.5; c.hdr[WRSN_H_NODES_SEQ] = take_seq(c);
            }
            break;
        case K_UNTIL:
            if (c.tid == 0) { c.hdr[WRSN_H_UNTIL_ON] = 0.0; c.bcast[2] = 1; }
            break;
        case K_SLOT: ev_slot(c, idx); break;
        case K_COND: ev_cond(c, idx); break;
        default:
            if (c.tid == 0) { c.hdr[WRSN_H_ERR] = 1.0; c.bcast[2] = 1; }
            break;
        }
        gsync(c);
        if (c.bcast[2]) break;
    }
    gsync(c);
}

/* ------------------------------------------------------------------ entry points (one environment) */

/* NetworkIO.makeNetwork + the t = 0 starts of Network.operate / update_reward / Node.operate */
WRSN_D void entry_init_network(Ctx &c, int with_reward) {
    for (int i = c.tid; i < c.Npad; i += c.G) {
        bool real = i < c.N;
        c.energy[i] = real ? c.par[WRSN_P_CAP] : 0.0;
        c.rr[i] = 0.0; c.cs[i] = 0.0; c.esend[i] = 0.0; c.logc[i] = 0.0;
        c.nbef[i] = 0; c.naft[i] = 0; c.level[i] = -1; c.parent[i] = -1;
        c.status[i] = real ? 1 : 0;
        c.logtick[i] = 0.0;
        for (int k = 0; k < WRSN_RING; k++) c.ring[(size_t)k * c.Npad + i] = 0.0;
    }
    for (int w = c.tid; w < c.Tw; w += c.G) {
        int bits = c.T - 32 * w; if (bits > 32) bits = 32;
        c.tact[w] = bits == 32 ? 0xffffffffu : ((1u << bits) - 1u);
    }
    for (int w = c.tid; w < (c.M > 0 ? c.M : 1) * c.W; w += c.G) c.conn[w] = 0u;
    for (int k = c.tid; k < WRSN_H_LEN; k += c.G) c.hdr[k] = 0.0;
    for (int k = c.tid; k < (c.M > 0 ? c.M : 1) * WRSN_MC_LEN; k += c.G) c.mc[k] = 0.0;
    for (int k = c.tid; k < c.n_slot * WRSN_PR_LEN; k += c.G) c.proc[k] = 0.0;
    gsync(c);
    for (int i = c.tid; i < c.N; i += c.G) check_status_node(c, i);   /* Node.__init__ :43 */
    if (c.tid == 0) {
        double *h = c.hdr;
        h[WRSN_H_ALIVE] = 1.0; h[WRSN_H_BFS_DIRTY] = 1.0; h[WRSN_H_CHAIN_DETACH] = -1.0;
        /* scheduling order at t = 0: Network.operate's timeout(0.1), update_reward's timeout(1.0), the nodes'
           timeout(0.5) (the process starts themselves are URGENT events at t = 0 and have all run) */
        h[WRSN_H_NET_ON] = 1.0; h[WRSN_H_NET_T] = 1.0 / 10.0; h[WRSN_H_NET_SEQ] = take_seq(c); h[WRSN_H_NET_STATE] = 1.0;
        if (with_reward) { h[WRSN_H_UR_ON] = 1.0; h[WRSN_H_UR_T] = 1.0; h[WRSN_H_UR_SEQ] = take_seq(c); }
        h[WRSN_H_NODES_T] = 0.5; h[WRSN_H_NODES_SEQ] = take_seq(c); h[WRSN_H_NODES_PHASE] = 1.0;
    }
    gsync(c);
}

/* env.run(until=t) */
WRSN_D void entry_run_until(Ctx &c, double at) {
    if (!(at > c.hdr[WRSN_H_NOW])) return;
    if (c.tid == 0) {
        c.hdr[WRSN_H_UNTIL_ON] = 1.0; c.hdr[WRSN_H_UNTIL_T] = at; c.hdr[WRSN_H_UNTIL_SEQ] = take_seq(c);
    }
    gsync(c);
    run_loop(c);
}

WRSN_D int scan_decider(Ctx &c) {                  /* WRSN.py:321-322 */
    for (int a = 0; a < c.M; a++) {
        const double *m = mc_of(c, a);
        if (euclid2(m[WRSN_MC_X], m[WRSN_MC_Y], m[WRSN_MC_CPA0], m[WRSN_MC_CPA1]) < c.par[WRSN_P_EPSENV] &&
            m[WRSN_MC_CPA2] == 0.0) return a;
    }
    return -1;
}

WRSN_D int new_slot(Ctx &c, int agent, double phy0, double phy1, double phy2) {   /* env.process(agent.operate_step(phy)) */
    int s = -1;
    for (int k = 0; k < c.n_slot; k++) if (slot_of(c, k)[WRSN_PR_USED] == 0.0) { s = k; break; }
    if (s < 0) { c.hdr[WRSN_H_ERR] = 3.0; return -1; }
    double *p = slot_of(c, s);
    for (int k = 0; k < WRSN_PR_LEN; k++) p[k] = 0.0;
    p[WRSN_PR_USED] = 1.0; p[WRSN_PR_CURRENT] = 1.0; p[WRSN_PR_AGENT] = agent;
    p[WRSN_PR_PHY0] = phy0; p[WRSN_PR_PHY1] = phy1; p[WRSN_PR_PHY2] = phy2;
    slot_sched(c, p, PC_OP_INIT, WRSN_URGENT, 0.0);
    return s;
}

struct ReqOut { int agent; int terminal; double reward, now, act[3], detail[2]; int flags; };

/* the rest of WRSN.reset after env.run(until=warm_up) (WRSN.py:44-83) */
WRSN_D void entry_reset_finish(Ctx &c, ReqOut *r) {
    if (c.tid == 0) {
        for (int a = 0; a < c.M; a++) {
            double *m = mc_of(c, a);
            for (int k = 0; k < WRSN_MC_LEN; k++) m[k] = 0.0;
            m[WRSN_MC_X] = c.par[WRSN_P_BSX]; m[WRSN_MC_Y] = c.par[WRSN_P_BSY];
            m[WRSN_MC_ENERGY] = c.par[WRSN_P_MC_CAP]; m[WRSN_MC_STATUS] = 1.0;
            mc_check_status(c, m);
            m[WRSN_MC_CPA0] = c.par[WRSN_P_BSX]; m[WRSN_MC_CPA1] = c.par[WRSN_P_BSY]; m[WRSN_MC_CPA2] = 0.0;
            m[WRSN_MC_SLOT] = -1.0;
        }
        for (int k = 0; k < c.n_slot * WRSN_PR_LEN; k++) c.proc[k] = 0.0;
        for (int w = 0; w < c.M * c.W; w++) c.conn[w] = 0u;
        c.hdr[WRSN_H_CHAIN_N] = 0.0; c.hdr[WRSN_H_CHAIN_DETACH] = -1.0; c.hdr[WRSN_H_HANG] = 0.0;
    }
    gsync(c);
    double fit = do_fitness(c, (double *)0);
    if (c.tid == 0) {
        c.hdr[WRSN_H_FIT_MIN] = fit;
        const double f0 = c.par[WRSN_P_F0], f1 = c.par[WRSN_P_F1], f2 = c.par[WRSN_P_F2], f3 = c.par[WRSN_P_F3];
        for (int a = 0; a < c.M; a++) {
            double *m = mc_of(c, a);
            m[WRSN_MC_ACT0] = (c.par[WRSN_P_BSX] - f0) / (f1 - f0);          /* down_mapping :86-88 */
            m[WRSN_MC_ACT1] = (c.par[WRSN_P_BSY] - f2) / (f3 - f2);
            m[WRSN_MC_ACT2] = 0.0;
            m[WRSN_MC_SLOT] = new_slot(c, a, m[WRSN_MC_CPA0], m[WRSN_MC_CPA1], m[WRSN_MC_CPA2]);
            m[WRSN_MC_PREVFIT] = fit; m[WRSN_MC_EXCL] = 0.0;
        }
        int id = scan_decider(c);
        r->agent = id; r->terminal = c.hdr[WRSN_H_ALIVE] == 1.0 ? 0 : 1; r->now = c.hdr[WRSN_H_NOW];
        r->reward = id >= 0 ? 0.0 : NAN; r->detail[0] = r->detail[1] = id >= 0 ? 0.0 : NAN;
        for (int k = 0; k < 3; k++) r->act[k] = id >= 0 ? mc_of(c, id)[WRSN_MC_ACT0 + k] : NAN;
        r->flags = c.hdr[WRSN_H_ERR] != 0.0 ? 2 : 0;
        if (id >= 0) c.hdr[WRSN_H_NDECISIONS] += 1.0;
    }
    gsync(c);
}

/* WRSN.step (WRSN.py:289-330) */
WRSN_D void entry_step(Ctx &c, int agent_id, const double *input_action, ReqOut *r) {
    if (c.tid == 0) {
        double *h = c.hdr;
        if (agent_id >= 0 && agent_id < c.M) {       /* :290-305 */
            double *m = mc_of(c, agent_id);
            double act[3];
            for (int k = 0; k < 3; k++) act[k] = fmin(fmax(input_action[k], 0.0), 1.0);   /* np.clip */
            for (int k = 0; k < 3; k++) m[WRSN_MC_ACT0 + k] = act[k];
            const double f0 = c.par[WRSN_P_F0], f1 = c.par[WRSN_P_F1], f2 = c.par[WRSN_P_F2], f3 = c.par[WRSN_P_F3];
            double phy0 = act[0] * (f1 - f0) + f0;                                        /* translate :95-98 */
            double phy1 = act[1] * (f3 - f2) + f2;
            double phy2 = c.par[WRSN_P_CTM] * act[2];
            int old = (int)m[WRSN_MC_SLOT];
            if (old >= 0) {
                double *po = slot_of(c, old);
                po[WRSN_PR_CURRENT] = 0.0;
                if (po[WRSN_PR_PROCESSED] != 0.0) po[WRSN_PR_USED] = 0.0;
            }
            m[WRSN_MC_SLOT] = new_slot(c, agent_id, phy0, phy1, phy2);
            m[WRSN_MC_PREVFIT] = h[WRSN_H_FIT_MIN];   /* the network has not moved since the last request */
            m[WRSN_MC_EXCL] = 0.0;
        }
        /* general_process = net_process | p_0 | p_1 ... over chargers with status != 0 (:307-310) */
        int n = 0;
        h[WRSN_H_CHAIN_DETACH] = -1.0;
        for (int a = 0; a < c.M; a++) {
            const double *m = mc_of(c, a);
            if (m[WRSN_MC_STATUS] == 0.0) continue;
            int s = (int)m[WRSN_MC_SLOT];
            h[WRSN_H_CHAIN_SLOT + n] = s; h[WRSN_H_COND_TRIG + n] = 0.0; h[WRSN_H_COND_PEND + n] = 0.0;
            h[WRSN_H_CHAIN_N] = n + 1;
            if (s >= 0 && slot_of(c, s)[WRSN_PR_PROCESSED] != 0.0) cond_check(c, n);   /* operand already processed */
            n++;
        }
        h[WRSN_H_CHAIN_N] = n;
        h[WRSN_H_HANG] = n == 0 ? 1.0 : 0.0;
        c.bcast[3] = n;
    }
    gsync(c);
    const int watched = c.bcast[3];
    gsync(c);
    if (watched > 0) run_loop(c);
    int id = -1;
    if (c.tid == 0) {
        if (watched == 0 || c.hdr[WRSN_H_ALIVE] == 0.0) id = -1;
        else { id = scan_decider(c); if (id < 0) id = -2; }
        c.bcast[4] = id;
    }
    gsync(c);
    id = c.bcast[4];
    gsync(c);
    double fit = 0.0;
    if (id >= 0) fit = do_fitness(c, (double *)0);   /* get_reward :222-227 */
    if (c.tid == 0) {
        r->now = c.hdr[WRSN_H_NOW];
        r->flags = (watched == 0 ? 1 : 0) | (c.hdr[WRSN_H_ERR] != 0.0 ? 2 : 0);
        r->agent = id;
        r->terminal = (c.hdr[WRSN_H_ALIVE] == 0.0) ? 1 : 0;
        if (id >= 0) {
            double *m = mc_of(c, id);
            c.hdr[WRSN_H_FIT_MIN] = fit;
            double term_all = fit - m[WRSN_MC_PREVFIT];
            double term_excl = m[WRSN_MC_EXCL] / c.par[WRSN_P_AVGNA];
            r->reward = (term_all * 0.8 + 0.2 * term_excl) / (c.par[WRSN_P_CTM] + c.par[WRSN_P_MTM]);
            r->detail[0] = term_all; r->detail[1] = term_excl;
            for (int k = 0; k < 3; k++) r->act[k] = m[WRSN_MC_ACT0 + k];
            c.hdr[WRSN_H_NDECISIONS] += 1.0;
        } else {
            r->reward = NAN; r->detail[0] = r->detail[1] = NAN;
            for (int k = 0; k < 3; k++) r->act[k] = NAN;
        }
    }
    gsync(c);
}

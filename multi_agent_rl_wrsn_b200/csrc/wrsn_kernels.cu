/*
 * wrsn_kernels.cu — sm_100a kernels and the C ABI (include/wrsn_b200.h) of the batched WRSN simulator.
 *
 * One CTA per environment.  A launch copies the environment's resident record from HBM into shared
 * memory with 16-byte vector loads, runs the engine (wrsn_engine.cuh) there, and writes it back.
 * There is no CPU path: every entry point launches a kernel or returns an error.
 */
#include <cuda_runtime.h>
#include <stdio.h>

#include "wrsn_engine.cuh"

static thread_local char g_err[512] = "";
#define WRSN_FAIL(...) do { snprintf(g_err, sizeof(g_err), __VA_ARGS__); return -1; } while (0)
#define WRSN_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) WRSN_FAIL("%s: %s", #x, cudaGetErrorString(e_)); } while (0)

enum { MODE_INIT = 0, MODE_RUN_UNTIL, MODE_RESET_FINISH, MODE_RESTORE_RESET, MODE_STEP, MODE_FITNESS, MODE_K_BFS,
       MODE_K_DRAIN, MODE_K_BOOK, MODE_K_REWARD };

struct KParams {
    wrsn_dims d;
    WrsnLayout L;
    const char *scen;
    const int32_t *scen_id;
    char *state;
    const char *snap;
    const uint8_t *mask;
    const double *t_until;
    const int32_t *agent_in;
    const double *action_in;
    wrsn_request req;
    double *fitness, *fit_min;
    int with_reward;
};

__device__ __forceinline__ void copy16(char *dst, const char *src, int64_t bytes, int tid, int G) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const int n = (int)(bytes >> 4);
    for (int i = tid; i < n; i += G) d[i] = s[i];
}

__device__ __forceinline__ void write_request(const wrsn_request &q, int b, const ReqOut &r) {
    if (q.agent_id) q.agent_id[b] = r.agent;
    if (q.terminal) q.terminal[b] = (uint8_t)r.terminal;
    if (q.reward) q.reward[b] = r.reward;
    if (q.now) q.now[b] = r.now;
    if (q.action) for (int k = 0; k < 3; k++) q.action[3 * b + k] = r.act[k];
    if (q.detail) { q.detail[2 * b] = r.detail[0]; q.detail[2 * b + 1] = r.detail[1]; }
    if (q.flags) q.flags[b] = r.flags;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_env(const KParams P) {
    extern __shared__ uint4 smem_u4[];
    char *smem = reinterpret_cast<char *>(smem_u4);
    const int b = blockIdx.x, tid = threadIdx.x, G = blockDim.x;
    if (P.mask && !P.mask[b]) return;
    char *row = P.state + (size_t)b * P.L.total;
    const char *scen_row = P.scen + (size_t)P.scen_id[b] * P.L.scen_total;
    Ctx c;
    ctx_bind(c, P.d, P.L, scen_row, row, smem, tid, G);
    if (MODE == MODE_RESTORE_RESET) {
        const char *src = P.snap + (size_t)P.scen_id[b] * P.L.total;
        copy16(row + P.L.resident, src + P.L.resident, P.L.total - P.L.resident, tid, G);
        copy16(smem, src, P.L.resident, tid, G);
    } else if (MODE != MODE_INIT) {
        copy16(smem, row, P.L.resident, tid, G);
    }
    for (int i = tid; i < c.Npad; i += G) c.own[i] = i < c.N ? (uint16_t)(c.tgt_ptr[i + 1] - c.tgt_ptr[i]) : (uint16_t)0;
    if (G == 32) __syncwarp(); else __syncthreads();

    ReqOut r;
    r.agent = -3; r.terminal = 0; r.reward = 0; r.now = 0; r.flags = 0;
    r.act[0] = r.act[1] = r.act[2] = 0; r.detail[0] = r.detail[1] = 0;
    switch (MODE) {
    case MODE_INIT: entry_init_network(c, P.with_reward); break;
    case MODE_RUN_UNTIL: entry_run_until(c, P.t_until[b]); break;
    case MODE_RESET_FINISH:
    case MODE_RESTORE_RESET: entry_reset_finish(c, &r); break;
    case MODE_STEP: entry_step(c, P.agent_in ? P.agent_in[b] : -1, P.action_in ? P.action_in + 3 * (size_t)b : nullptr, &r); break;
    case MODE_FITNESS: {
        double mn = do_fitness(c, P.fitness ? P.fitness + (size_t)b * P.d.T : nullptr);
        if (tid == 0 && P.fit_min) P.fit_min[b] = mn;
        break;
    }
    case MODE_K_BFS: do_bfs(c); break;
    case MODE_K_DRAIN: ev_nodes_drain(c); break;
    case MODE_K_BOOK: ev_nodes_book(c); break;
    case MODE_K_REWARD: ev_update_reward(c); break;
    }
    if (G == 32) __syncwarp(); else __syncthreads();
    if (MODE != MODE_FITNESS) copy16(row, smem, P.L.resident, tid, G);
    if ((MODE == MODE_RESET_FINISH || MODE == MODE_RESTORE_RESET || MODE == MODE_STEP) && tid == 0) write_request(P.req, b, r);
}

/* ------------------------------------------------------------------ WRSN.get_state (rl_env/WRSN.py:130-186)
 * 4 x S x S map per environment.  Every source (alive node, charger) contributes w * G(x - x0; hX) * G(y - y0; hY)
 * with G(u; h) = exp(u^2 / (-2 h^2)) — separable, so a chunk of sources is expanded into two S-vectors each in
 * shared memory (2 S exponentials per source instead of S^2) and every thread accumulates its own output cells
 * over the sources IN SOURCE ORDER, in fp64, exactly as the reference's `map += pdf` does. */
#define OBS_THREADS 256
#define OBS_CHUNK 16
#define OBS_EPT 40            /* output cells per thread per pass: 256 * 40 >= 100 * 100 */

struct ObsSrc { double x0, y0, hx, hy, w; int mode; };     /* mode 0: (w*gx)*gy ; 1: gx*gy*w/mtm */

template <typename OutT>
__global__ void __launch_bounds__(OBS_THREADS) k_observe(const KParams P, const int32_t *agent_id, OutT *obs) {
    extern __shared__ uint4 smem_u4[];
    double *gx = reinterpret_cast<double *>(smem_u4);                 /* [OBS_CHUNK][S] */
    const int S = P.d.S, N = P.d.N, M = P.d.M;
    double *gy = gx + OBS_CHUNK * S;                                  /* [OBS_CHUNK][S] */
    __shared__ ObsSrc src[OBS_CHUNK];
    __shared__ int nsrc_s;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int ag = agent_id[b];
    if (ag < 0) return;
    const char *row = P.state + (size_t)b * P.L.total;
    const char *scen_row = P.scen + (size_t)P.scen_id[b] * P.L.scen_total;
    const double *par = (const double *)(scen_row + P.L.soff[WRSN_S_PAR]);
    const double *nx = (const double *)(scen_row + P.L.soff[WRSN_S_NX]);
    const double *ny = (const double *)(scen_row + P.L.soff[WRSN_S_NY]);
    const double *energy = (const double *)(row + P.L.off[WRSN_F_ENERGY]);
    const double *cs = (const double *)(row + P.L.off[WRSN_F_CS]);
    const uint8_t *status = (const uint8_t *)(row + P.L.off[WRSN_F_STATUS]);
    const double *mc = (const double *)(row + P.L.off[WRSN_F_MC]);
    const double f0 = par[WRSN_P_F0], f1 = par[WRSN_P_F1], f2 = par[WRSN_P_F2], f3 = par[WRSN_P_F3];
    const double Wd = f1 - f0, Hd = f3 - f2;
    const double unit = 1.0 / (double)S, start = unit / 2.0, delta = (start + unit) - start;   /* np.arange(unit/2, 1.0, unit) */
    const double R = par[WRSN_P_MC_R];
    const double *me = mc + (size_t)ag * WRSN_MC_LEN;
    const int SS = S * S;
    OutT *out = obs + (size_t)b * 4 * SS;

    for (int base = 0; base < SS; base += OBS_THREADS * OBS_EPT) {
        for (int ch = 0; ch < 4; ch++) {
            double acc[OBS_EPT];
#pragma unroll
            for (int k = 0; k < OBS_EPT; k++) acc[k] = 0.0;
            const int total = ch == 0 ? N : (ch == 1 ? 1 : M);
            for (int s0 = 0; s0 < total; s0 += OBS_CHUNK) {
                __syncthreads();
                if (tid == 0) {                      /* gather this chunk's sources, in id order */
                    int n = 0;
                    for (int s = s0; s < total && s < s0 + OBS_CHUNK; s++) {
                        ObsSrc q; q.mode = 0;
                        if (ch == 0) {
                            if (status[s] == 0) continue;
                            q.x0 = (nx[s] - f0) / Wd; q.y0 = (ny[s] - f2) / Hd; q.hx = R / Wd; q.hy = R / Hd;
                            q.w = (cs[s] / par[WRSN_P_MC_AB2]) / ((energy[s] - par[WRSN_P_THR]) / par[WRSN_P_CAPMTHR]);
                        } else if (ch == 1) {
                            double tmp = fmin(Hd, Wd);
                            q.x0 = (me[WRSN_MC_X] - f0) / Wd; q.y0 = (me[WRSN_MC_Y] - f2) / Hd;
                            q.hx = 0.5 * tmp / Wd; q.hy = 0.5 * tmp / Hd;
                            q.w = me[WRSN_MC_ENERGY] / par[WRSN_P_MC_CAP];
                        } else {
                            if (s == ag) continue;
                            const double *an = mc + (size_t)s * WRSN_MC_LEN;
                            bool charging = an[WRSN_MC_TYPE] != 0.0;
                            if (ch == 2 ? !charging : charging) continue;
                            q.x0 = (an[WRSN_MC_CPA0] - f0) / Wd; q.y0 = (an[WRSN_MC_CPA1] - f2) / Hd; q.hx = R / Wd; q.hy = R / Hd;
                            if (ch == 2) q.w = an[WRSN_MC_CPA2] / par[WRSN_P_CTM];
                            else {                   /* the observer's destination y */
                                double dx = an[WRSN_MC_X] - an[WRSN_MC_CPA0], dy = an[WRSN_MC_Y] - me[WRSN_MC_CPA1];
                                q.w = sqrt(dx * dx + dy * dy) / par[WRSN_P_MC_V];
                                q.mode = 1;
                            }
                        }
                        src[n++] = q;
                    }
                    nsrc_s = n;
                }
                __syncthreads();
                const int n = nsrc_s;
                for (int k = tid; k < n * S; k += OBS_THREADS) {
                    int q = k / S, i = k - q * S;
                    double cc = start + (double)i * delta;
                    double ux = cc - src[q].x0, uy = cc - src[q].y0;
                    double ex = exp(ux * ux / (-2.0 * (src[q].hx * src[q].hx)));
                    double ey = exp(uy * uy / (-2.0 * (src[q].hy * src[q].hy)));
                    gx[k] = src[q].mode == 0 ? src[q].w * ex : ex;
                    gy[k] = ey;
                }
                __syncthreads();
                for (int q = 0; q < n; q++) {
                    const double *gxq = gx + q * S, *gyq = gy + q * S;
                    if (src[q].mode == 0) {
#pragma unroll
                        for (int k = 0; k < OBS_EPT; k++) {
                            int o = base + tid + k * OBS_THREADS;
                            if (o < SS) { int i = o / S, j = o - i * S; acc[k] += gxq[i] * gyq[j]; }
                        }
                    } else {
                        const double w = src[q].w, mtm = par[WRSN_P_MTM];
#pragma unroll
                        for (int k = 0; k < OBS_EPT; k++) {
                            int o = base + tid + k * OBS_THREADS;
                            if (o < SS) { int i = o / S, j = o - i * S; acc[k] += gxq[i] * gyq[j] * w / mtm; }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < OBS_EPT; k++) {
                int o = base + tid + k * OBS_THREADS;
                if (o < SS) out[(size_t)ch * SS + o] = (OutT)acc[k];
            }
        }
    }
}

/* ------------------------------------------------------------------ host side of the C ABI */
static int check_dims(const wrsn_dims *d) {
    if (!d) WRSN_FAIL("dims is NULL");
    if (d->N <= 0 || d->N > 32000 || d->T < 0 || d->T > 65000 || d->M < 0 || d->M > WRSN_MAX_MC || d->B <= 0 || d->S <= 0)
        WRSN_FAIL("bad dims N=%d T=%d M=%d B=%d S=%d", d->N, d->T, d->M, d->B, d->S);
    if (d->Npad < d->N || (d->Npad & 15) || d->state_bytes <= 0) WRSN_FAIL("dims not finalized (call wrsn_dims_finalize)");
    return 0;
}

extern "C" {

const char *wrsn_last_error(void) { return g_err; }
int wrsn_abi_version(void) { return WRSN_ABI_VERSION; }
int wrsn_field_count(int which) {
    switch (which) {
    case 0: return WRSN_P_LEN; case 1: return WRSN_H_LEN; case 2: return WRSN_MC_LEN; case 3: return WRSN_PR_LEN;
    case 4: return WRSN_F_COUNT; case 5: return WRSN_S_COUNT; default: return -1;
    }
}

int wrsn_dims_finalize(wrsn_dims *d) {
    if (!d) WRSN_FAIL("dims is NULL");
    d->Npad = (d->N + 15) & ~15;
    d->W = (d->N + 31) / 32;
    d->Tw = (d->T + 31) / 32; if (d->Tw < 1) d->Tw = 1;
    d->n_slot = 2 * d->M + 2;
    if (d->Emax < 1) d->Emax = 1;
    if (d->TEmax < 1) d->TEmax = 1;
    if (d->threads <= 0) {
        int per = (d->N + 3) / 4;                    /* about four nodes per thread */
        int t = 32; while (t < per && t < 256) t *= 2;
        d->threads = t;
    }
    if (d->threads % 32 || d->threads > 256) WRSN_FAIL("threads must be a multiple of 32, at most 256");
    WrsnLayout L;
    wrsn_make_layout(d, &L);
    d->state_bytes = (int32_t)L.total; d->state_resident_bytes = (int32_t)L.resident;
    d->scen_bytes = (int32_t)L.scen_total; d->smem_bytes = (int32_t)L.smem_total;
    if (L.smem_total > 227 * 1024) WRSN_FAIL("environment does not fit shared memory (%lld bytes)", (long long)L.smem_total);
    return 0;
}

int wrsn_state_layout(const wrsn_dims *d, int64_t *offsets) {
    if (check_dims(d)) return -1;
    WrsnLayout L; wrsn_make_layout(d, &L);
    for (int k = 0; k < WRSN_F_COUNT; k++) offsets[k] = L.off[k];
    return 0;
}
int wrsn_scen_layout(const wrsn_dims *d, int64_t *offsets) {
    if (check_dims(d)) return -1;
    WrsnLayout L; wrsn_make_layout(d, &L);
    for (int k = 0; k < WRSN_S_COUNT; k++) offsets[k] = L.soff[k];
    return 0;
}

int wrsn_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); snprintf(g_err, sizeof(g_err), "no CUDA device"); return 0; }
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    if (p.major != 10) { snprintf(g_err, sizeof(g_err), "device is sm_%d%d, this library is built for sm_100a only", p.major, p.minor); return 0; }
    return 1;
}

}  /* extern "C" */

template <int MODE>
static int launch_env(KParams &P, void *stream) {
    if (check_dims(&P.d)) return -1;
    wrsn_make_layout(&P.d, &P.L);
    if (!P.scen || !P.scen_id || !P.state) WRSN_FAIL("scen / scen_id / state must not be NULL");
    static int64_t attr_bytes = 48 * 1024;           /* per template instance; the opt-in limit only ever grows */
    if (P.L.smem_total > attr_bytes) {
        WRSN_CUDA(cudaFuncSetAttribute(k_env<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.L.smem_total));
        attr_bytes = P.L.smem_total;
    }
    k_env<MODE><<<P.d.B, P.d.threads, (size_t)P.L.smem_total, (cudaStream_t)stream>>>(P);
    WRSN_CUDA(cudaGetLastError());
    return 0;
}

static KParams base_params(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state, const uint8_t *mask) {
    KParams P;
    memset(&P, 0, sizeof(P));
    if (d) P.d = *d;
    P.scen = (const char *)scen; P.scen_id = scen_id; P.state = (char *)state; P.mask = mask;
    return P;
}

extern "C" {

int wrsn_init_network(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state,
                      const uint8_t *env_mask, int with_reward_process, void *stream) {
    KParams P = base_params(d, scen, scen_id, state, env_mask);
    P.with_reward = with_reward_process;
    return launch_env<MODE_INIT>(P, stream);
}

int wrsn_run_until(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state,
                   const uint8_t *env_mask, const double *t_until, void *stream) {
    if (!t_until) WRSN_FAIL("t_until is NULL");
    KParams P = base_params(d, scen, scen_id, state, env_mask);
    P.t_until = t_until;
    return launch_env<MODE_RUN_UNTIL>(P, stream);
}

int wrsn_reset_finish(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state,
                      const uint8_t *env_mask, wrsn_request *req, void *stream) {
    if (!req) WRSN_FAIL("req is NULL");
    KParams P = base_params(d, scen, scen_id, state, env_mask);
    P.req = *req;
    return launch_env<MODE_RESET_FINISH>(P, stream);
}

int wrsn_reset_from_snapshot(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state,
                             const void *snap, const uint8_t *env_mask, wrsn_request *req, void *stream) {
    if (!req || !snap) WRSN_FAIL("req / snap is NULL");
    KParams P = base_params(d, scen, scen_id, state, env_mask);
    P.req = *req; P.snap = (const char *)snap;
    return launch_env<MODE_RESTORE_RESET>(P, stream);
}

int wrsn_step(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state,
              const uint8_t *env_mask, const int32_t *agent_id_in, const double *action_in,
              wrsn_request *req, void *stream) {
    if (!req) WRSN_FAIL("req is NULL");
    if (agent_id_in && !action_in) WRSN_FAIL("action_in is NULL");
    KParams P = base_params(d, scen, scen_id, state, env_mask);
    P.agent_in = agent_id_in; P.action_in = action_in; P.req = *req;
    return launch_env<MODE_STEP>(P, stream);
}

int wrsn_fitness(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state,
                 double *fitness, double *fit_min, void *stream) {
    KParams P = base_params(d, scen, scen_id, state, nullptr);
    P.fitness = fitness; P.fit_min = fit_min;
    return launch_env<MODE_FITNESS>(P, stream);
}

int wrsn_observe(const wrsn_dims *d, const void *scen, const int32_t *scen_id, const void *state,
                 const int32_t *agent_id, void *obs, int obs_f64, void *stream) {
    if (check_dims(d)) return -1;
    if (!agent_id || !obs || !scen || !scen_id || !state) WRSN_FAIL("NULL argument");
    KParams P = base_params(d, scen, scen_id, const_cast<void *>(state), nullptr);
    wrsn_make_layout(&P.d, &P.L);
    size_t smem = sizeof(double) * 2 * OBS_CHUNK * (size_t)d->S;
    if (smem > 200 * 1024) WRSN_FAIL("map_size too large");
    if (obs_f64) {
        if (smem > 48 * 1024) WRSN_CUDA(cudaFuncSetAttribute(k_observe<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_observe<double><<<d->B, OBS_THREADS, smem, (cudaStream_t)stream>>>(P, agent_id, (double *)obs);
    } else {
        if (smem > 48 * 1024) WRSN_CUDA(cudaFuncSetAttribute(k_observe<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_observe<float><<<d->B, OBS_THREADS, smem, (cudaStream_t)stream>>>(P, agent_id, (float *)obs);
    }
    WRSN_CUDA(cudaGetLastError());
    return 0;
}

int wrsn_k_bfs(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state, void *stream) {
    KParams P = base_params(d, scen, scen_id, state, nullptr);
    return launch_env<MODE_K_BFS>(P, stream);
}
int wrsn_k_drain(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state, void *stream) {
    KParams P = base_params(d, scen, scen_id, state, nullptr);
    return launch_env<MODE_K_DRAIN>(P, stream);
}
int wrsn_k_bookkeep(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state, void *stream) {
    KParams P = base_params(d, scen, scen_id, state, nullptr);
    return launch_env<MODE_K_BOOK>(P, stream);
}
int wrsn_k_reward(const wrsn_dims *d, const void *scen, const int32_t *scen_id, void *state, void *stream) {
    KParams P = base_params(d, scen, scen_id, state, nullptr);
    return launch_env<MODE_K_REWARD>(P, stream);
}

}  /* extern "C" */

/*
 * wrsn_layout.h — byte layout of one environment record / one scenario record / the shared-memory image
 * (shared by the host side of the C ABI, the kernels and the tests' host emulation).
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "wrsn_b200.h"

#if defined(WRSN_HOST_EMU)
#define WRSN_HD static inline
#else
#define WRSN_HD __host__ __device__ static inline
#endif

#define WRSN_SPEC_MAX 16                            /* charged (or otherwise irregular) nodes a whole-cycle batch handles by table */
#define WRSN_PAIR_MAX 32                            /* (charger, node) pairs the incentive sums of a batch handle by list */
#define WRSN_SPEC_LEN 6                             /* D1, D2, H, lo guard, hi guard, node id */

/* ------------------------------------------------------------------ layouts */
struct WrsnLayout {
    int64_t off[WRSN_F_COUNT];
    int64_t resident, total;                        /* bytes mirrored in shared memory / bytes per record */
    int64_t s_own, s_scr0, s_scr1, s_bcast, s_red, s_par, s_spec, s_exptab, s_pairs, smem_total;
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
    L->off[WRSN_F_SCRATCH] = o; o += wrsn_a16(17 * Np);
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
    L->s_par = s; s += 8 * WRSN_P_LEN;             /* scenario constants: copied next to the state, never re-read from HBM */
    L->s_spec = s; s += 8 * WRSN_SPEC_LEN * WRSN_SPEC_MAX;   /* per-batch table of the charged nodes (wrsn_engine.cuh: reward_cycles) */
    L->s_exptab = s; s += 8 * 64;                  /* 2^(j/64): table of the reward path's exponential */
    L->s_pairs = s; s += 8 * WRSN_PAIR_MAX + 4 * 2 * WRSN_PAIR_MAX + 4 * 2 * WRSN_MAX_MC;   /* (charging charger, connected alive node) pairs of the current batch:
                                                      term double[PAIR_MAX]; {charger, node} int[PAIR_MAX][2]; per charger {first, end} int[MAX_MC][2] */
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
    o = wrsn_a16(o);
    L->soff[WRSN_S_OBS_GX] = o; o += 4 * Np * (int64_t)d->obs_pitch;
    L->soff[WRSN_S_OBS_GY] = o; o += 4 * Np * (int64_t)d->obs_pitch;
    L->scen_total = wrsn_a16(o);
}


/* request record of one environment, as the engine's entry points produce it */
struct ReqOut { int agent; int terminal; double reward, now, act[3], detail[2]; int flags; };

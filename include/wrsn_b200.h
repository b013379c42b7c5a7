/*
 * wrsn_b200.h — C ABI of the B200-native batched WRSN simulator (libwrsn_b200.so).
 *
 * The reference has no FFI layer: its boundary for this hot path is the Python class
 * rl_env.WRSN.WRSN (reset :41, step :289, get_state :130, get_network_fitness :188,
 * get_reward :222, update_reward :100) over physical_env/{network,mc}.  The entry points
 * below are what a ctypes binding of that class binds instead of running SimPy:
 * plain pointers and sizes, device pointers owned by the caller (PyTorch tensors are only
 * the allocator), a CUDA stream handle, int return codes (0 = ok) and wrsn_last_error().
 * INTEGRATION.md shows the reference-side stub.
 *
 * All state is struct-of-arrays in HBM:  per-environment node rows [B][Npad], per-scenario
 * static graph rows [n_scen][...].  Records that only the per-environment leader thread
 * touches (event clock, charger records, charger process slots) are rows of doubles whose
 * field indices are the enums below, so a host can read them without knowing a C layout.
 */
#ifndef WRSN_B200_H
#define WRSN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WRSN_MAX_MC 16          /* chargers per environment */
#define WRSN_RING 10            /* Node.operate keeps the last 10 per-second consumptions (Node.py:70-77) */

/* ---- per-scenario scalar row  par[n_scen][WRSN_P_LEN]  (filled by the host loader with the
 *      reference's own Python arithmetic so thresholds / constants are bit-identical) ---- */
enum {
    WRSN_P_CAP = 0, WRSN_P_THR, WRSN_P_ERECV,          /* node capacity, threshold, er*package_size */
    WRSN_P_BSX, WRSN_P_BSY,
    WRSN_P_F0, WRSN_P_F1, WRSN_P_F2, WRSN_P_F3,        /* Network.frame (Network.py:16-26) */
    WRSN_P_MAXTIME, WRSN_P_WARMUP,
    WRSN_P_MTM, WRSN_P_CTM, WRSN_P_AVGNA,              /* moving_time_max, charging_time_max, avg_nodes_agent (WRSN.py:50-52) */
    WRSN_P_MC_CAP, WRSN_P_MC_THR, WRSN_P_MC_V, WRSN_P_MC_PM, WRSN_P_MC_R, WRSN_P_MC_ALPHA, WRSN_P_MC_BETA,
    WRSN_P_MC_EPS, WRSN_P_MC_AB2,                      /* alpha / beta**2 */
    WRSN_P_MC_CAP200, WRSN_P_MC_PMV,                   /* capacity / 200.0 ; pm * velocity */
    WRSN_P_DENX1, WRSN_P_DENY1, WRSN_P_DENX2, WRSN_P_DENY2,  /* -2*hX**2 terms of get_state (WRSN.py:16-18,144-155) */
    WRSN_P_EPSENV,                                     /* WRSN.epsilon = 1e-9 */
    WRSN_P_CAPMTHR,                                    /* capacity - threshold */
    WRSN_P_LEN = 40
};

/* ---- per-environment event clock  hdr[B][WRSN_H_LEN] ---- */
enum {
    WRSN_H_NOW = 0, WRSN_H_SEQ,
    WRSN_H_NET_T, WRSN_H_NET_SEQ, WRSN_H_NET_STATE,    /* Network.operate: 1 next=setLevels, 2 next=exit check, 3 finished */
    WRSN_H_UR_T, WRSN_H_UR_SEQ, WRSN_H_UR_ON,          /* WRSN.update_reward */
    WRSN_H_NODES_T, WRSN_H_NODES_SEQ, WRSN_H_NODES_PHASE, /* Node.operate block: 1 next=k+0.5 drain, 2 next=k+1.0 bookkeeping */
    WRSN_H_ALIVE, WRSN_H_BFS_DIRTY, WRSN_H_TREE_DIRTY, WRSN_H_LOG_LEN, WRSN_H_LOG_HEAD,
    WRSN_H_UNTIL_T, WRSN_H_UNTIL_SEQ, WRSN_H_UNTIL_ON,
    WRSN_H_CHAIN_N, WRSN_H_ERR, WRSN_H_HANG,
    WRSN_H_NTICKS, WRSN_H_NEVENTS, WRSN_H_NDEATH_TICKS, WRSN_H_NBFS,
    WRSN_H_CHAIN_AGENT = 32,                           /* [WRSN_MAX_MC] agent id of chain member j */
    WRSN_H_CHAIN_TRIG = WRSN_H_CHAIN_AGENT + WRSN_MAX_MC,
    WRSN_H_COND_ON = WRSN_H_CHAIN_TRIG + WRSN_MAX_MC,
    WRSN_H_COND_T = WRSN_H_COND_ON + WRSN_MAX_MC,
    WRSN_H_COND_SEQ = WRSN_H_COND_T + WRSN_MAX_MC,
    WRSN_H_LEN = WRSN_H_COND_SEQ + WRSN_MAX_MC         /* 112 */
};

/* ---- charger record  mc[B][M][WRSN_MC_LEN]  (MobileCharger.py:6-32 + WRSN per-agent lists) ---- */
enum {
    WRSN_MC_X = 0, WRSN_MC_Y, WRSN_MC_ENERGY, WRSN_MC_STATUS,
    WRSN_MC_CPA0, WRSN_MC_CPA1, WRSN_MC_CPA2,          /* cur_phy_action */
    WRSN_MC_TYPE,                                      /* 0 "moving", 1 "charging" */
    WRSN_MC_RATE, WRSN_MC_CHTIME,
    WRSN_MC_EXCL, WRSN_MC_PREVFIT,                     /* agents_exclusive_reward, min(agents_prev_fitness) */
    WRSN_MC_ACT0, WRSN_MC_ACT1, WRSN_MC_ACT2,          /* agents_action (clipped) */
    WRSN_MC_LEN = 16
};

/* ---- charger process slot  proc[B][M][2][WRSN_PR_LEN]  (slot 0 = agents_process[id], slot 1 = a
 *      still-running predecessor, e.g. the reset-time process of agent 0) ---- */
enum {
    WRSN_PR_ACTIVE = 0, WRSN_PR_DONE, WRSN_PR_STATE, WRSN_PR_T, WRSN_PR_PRIO, WRSN_PR_SEQ,
    WRSN_PR_PHY0, WRSN_PR_PHY1, WRSN_PR_PHY2, WRSN_PR_STAGE,
    WRSN_PR_DESTX, WRSN_PR_DESTY, WRSN_PR_MT, WRSN_PR_VX, WRSN_PR_VY, WRSN_PR_TOTAL, WRSN_PR_SPAN,
    WRSN_PR_SVX, WRSN_PR_SVY, WRSN_PR_CHTMP, WRSN_PR_CHSPAN,
    WRSN_PR_LEN = 24
};

typedef struct wrsn_dims {
    int32_t B;        /* environments in this shard */
    int32_t N, T, M;  /* nodes, targets, chargers (same for every scenario of the batch) */
    int32_t S;        /* map_size */
    int32_t Npad;     /* row pitch of node arrays (multiple of 4) */
    int32_t Tpad;     /* row pitch of target arrays */
    int32_t Emax;     /* row pitch of neighbour CSR payloads */
    int32_t TEmax;    /* row pitch of node->target CSR payload */
    int32_t W;        /* 32-bit words per node bitmask = ceil(N/32) */
    int32_t n_scen;   /* number of distinct scenarios */
    int32_t threads;  /* CTA size for per-environment kernels (one CTA per environment) */
} wrsn_dims;

/* static graph, device pointers, one row per scenario (Node.probe_neighbors :80, probe_targets :86,
 * BaseStation.probe_neighbors :20, send_package's e_send :107-115 precomputed per edge) */
typedef struct wrsn_static {
    const double *node_x, *node_y;      /* [n_scen][Npad] */
    const int32_t *nbr_ptr;             /* [n_scen][Npad+1] */
    const int32_t *nbr_idx;             /* [n_scen][Emax]  neighbour ids in id order */
    const double *nbr_dist;             /* [n_scen][Emax]  euclidean(node, neighbour) */
    const double *nbr_esend;            /* [n_scen][Emax]  e_send for that hop */
    const int32_t *tgt_ptr;             /* [n_scen][Npad+1] */
    const int32_t *tgt_idx;             /* [n_scen][TEmax] covered target ids in id order */
    const uint8_t *direct;              /* [n_scen][Npad]  d(node, BS) <= com_range */
    const double *bs_esend;             /* [n_scen][Npad]  e_send straight to the base station */
    const double *par;                  /* [n_scen][WRSN_P_LEN] */
} wrsn_static;

/* dynamic state, device pointers, one row per environment */
typedef struct wrsn_state {
    const int32_t *scen_id;             /* [B] */
    double *energy, *cs, *rr, *log_energy;   /* [B][Npad]  Node.energy / energyCS / energyRR / log_energy */
    double *ring;                       /* [B][WRSN_RING][Npad]  Node.log */
    uint8_t *status;                    /* [B][Npad] */
    int32_t *level, *parent;            /* [B][Npad]  Node.level; receiver (-2 base station, -1 none) */
    int32_t *nbef, *naft;               /* [B][Npad]  relayed packets per tick from lower / higher source ids */
    double *esend, *logc;               /* [B][Npad]  e_send to the current receiver; cached per-tick log_energy */
    uint8_t *targets_active;            /* [B][Tpad] */
    double *scratch;                    /* [B][2*max(Npad,Tpad)] */
    uint32_t *nearmask;                 /* [B][W] */
    double *hdr;                        /* [B][WRSN_H_LEN] */
    double *mc;                         /* [B][M][WRSN_MC_LEN] */
    double *proc;                       /* [B][M][2][WRSN_PR_LEN] */
    uint32_t *conn;                     /* [B][M][W]  MobileCharger.connected_nodes as a bitmask */
} wrsn_state;

/* request record written by reset / step, one row per environment */
typedef struct wrsn_request {
    int32_t *agent_id;                  /* [B]  -1 = None (terminal), -2 = implicit None (Q7) */
    uint8_t *terminal;                  /* [B] */
    double *reward;                     /* [B] */
    double *now;                        /* [B]  env.now */
    double *action;                     /* [B][3] agents_action[agent_id] */
    double *detail;                     /* [B][2] term_all, term_exclusive of get_reward (WRSN.py:225-226) */
} wrsn_request;

const char *wrsn_last_error(void);
int wrsn_abi_version(void);
int wrsn_sizeof_dims(void);
int wrsn_field_count(int which);        /* 0 par, 1 hdr, 2 mc, 3 proc */

/* NetworkIO.makeNetwork + Network.operate start: initialise node / clock state at t = 0.
 * with_reward_process != 0 also starts WRSN.update_reward (WRSN.py:43). `env_mask` (may be NULL) selects rows. */
int wrsn_init_network(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, const uint8_t *env_mask,
                      int with_reward_process, void *stream);
/* env.run(until=t) (WRSN.py:53): advance every selected environment to simulated time t_until[b]. */
int wrsn_run_until(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, const uint8_t *env_mask,
                   const double *t_until, void *stream);
/* remainder of WRSN.reset after the warm-up (:44-83): chargers at the base station, fitness, reset-time
 * charger processes, first request. */
int wrsn_reset_finish(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, const uint8_t *env_mask,
                      wrsn_request *req, void *stream);
/* WRSN.step (:289-330): agent_id_in[b] < 0 = "no new action" (agent_id None). */
int wrsn_step(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, const int32_t *agent_id_in,
              const double *action_in, wrsn_request *req, void *stream);
/* WRSN.get_state (:130-186) for agent agent_id[b] of every environment with agent_id[b] >= 0, written to
 * obs[b][4][S][S] as float (obs_f64 == 0) or double. */
int wrsn_observe(const wrsn_dims *d, const wrsn_static *st, const wrsn_state *s, const int32_t *agent_id,
                 void *obs, int obs_f64, void *stream);
/* WRSN.get_network_fitness (:188-220): per-target values fitness[B][Tpad] and their minimum. */
int wrsn_fitness(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, double *fitness, double *fit_min,
                 void *stream);

/* standalone per-tick kernels (same device functions the fused step uses; unit parity + ncu evidence) */
int wrsn_k_bfs(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, void *stream);       /* Network.setLevels + receivers + relay counts */
int wrsn_k_drain(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, void *stream);     /* Node.operate k+0.5 tick */
int wrsn_k_bookkeep(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, void *stream);  /* Node.operate k+1.0 tick */
int wrsn_k_charge(const wrsn_dims *d, const wrsn_static *st, wrsn_state *s, int connect, void *stream); /* charger_(dis)connection for every charging charger */

#ifdef __cplusplus
}
#endif
#endif

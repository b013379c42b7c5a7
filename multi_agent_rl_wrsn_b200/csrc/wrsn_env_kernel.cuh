/* wrsn_env_kernel.cuh — the one-CTA-per-environment kernel; included once per group-size specialisation (see
 * wrsn_engine.cuh), inside the same namespace. */
template <int MODE>
__global__ void __launch_bounds__(WRSN_GFIX ? WRSN_GFIX : 256, WRSN_GFIX ? 16 : 1) k_env(const KParams P) {
    char *smem = reinterpret_cast<char *>(wrsn_smem_u4);
    const int b = blockIdx.x, tid = threadIdx.x, G = blockDim.x;
    if (P.mask && !P.mask[b]) return;
    if (P.mask_mode == 1 && P.req.agent_id[b] < 0 && P.req.agent_id[b] != -4) return;     /* (-4: a step in flight continues) */
    if (P.mask_mode == 2 && (P.req.agent_id[b] >= 0 || P.req.agent_id[b] == -4)) return;
    if (P.resume_only && P.req.agent_id[b] != -4) return;
    char *row = P.state + (size_t)b * P.L.total;
    const bool split = P.d.step_budget > 0 && P.d.step_rounds > 0;
    if (MODE == MODE_STEP && split && P.req.agent_id && P.req.agent_id[b] == -4 &&
        reinterpret_cast<const double *>(row + P.L.off[WRSN_F_HDR])[WRSN_H_INFLIGHT] == 2.0) return;   /* waits for the batch kernel */
    if (MODE == MODE_STEP_BATCH && (P.req.agent_id[b] != -4 ||
        reinterpret_cast<const double *>(row + P.L.off[WRSN_F_HDR])[WRSN_H_INFLIGHT] != 2.0)) return;
    const char *scen_row = P.scen + (size_t)P.scen_id[b] * P.L.scen_total;
    Ctx c;
    ctx_bind(c, P.d, P.L, scen_row, row, tid, G);
    if (MODE == MODE_RESTORE_RESET) {
        const char *src = P.snap + (size_t)P.scen_id[b] * P.L.total;
        copy16(row + P.L.resident, src + P.L.resident, P.L.total - P.L.resident, tid, G);
        copy16(smem, src, P.L.resident, tid, G);
    } else if (MODE != MODE_INIT) {
        copy16(smem, row, P.L.resident, tid, G);
    } else {                                         /* alignment gaps between the fields: defined bytes in the record */
        uint4 *z = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < (int)(P.L.resident >> 4); i += G) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < c.Npad; i += G) c.own[i] = i < c.N ? (uint16_t)(c.tgt_ptr[i + 1] - c.tgt_ptr[i]) : (uint16_t)0;
    gsync(c);

    const double now_before = c.hdr[WRSN_H_NOW];
#if defined(WRSN_PROF)
    const long long prof_start = clock64();
    if (MODE == MODE_STEP) { gsync(c); if (tid == 0) for (int q = WRSN_H_PROF0; q <= WRSN_H_PROF4; q++) c.hdr[q] = 0.0; gsync(c); }
#endif
    ReqOut r;
    r.agent = -3; r.terminal = 0; r.reward = 0; r.now = 0; r.flags = 0;
    r.act[0] = r.act[1] = r.act[2] = 0; r.detail[0] = r.detail[1] = 0;
    switch (MODE) {
    case MODE_INIT: entry_init_network(c, P.with_reward); break;
    case MODE_RUN_UNTIL: entry_run_until(c, P.t_until[b]); break;
    case MODE_RESET_FINISH:
    case MODE_RESTORE_RESET: entry_reset_finish(c, &r); break;
    case MODE_STEP: entry_step(c, P.agent_in ? P.agent_in[b] : -1, P.action_in ? P.action_in + 3 * (size_t)b : nullptr, &r, P.d.step_budget, split ? 1 : 0); break;
    case MODE_STEP_BATCH: entry_batches(c, &r, P.d.step_budget); break;
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
    gsync(c);
#if defined(WRSN_PROF)
    if (MODE == MODE_STEP) { if (tid == 0) c.hdr[WRSN_H_PROF0] = (double)(clock64() - prof_start); gsync(c); }
#endif
    if (MODE != MODE_FITNESS) copy16(row, smem, P.L.resident, tid, G);
    if ((MODE == MODE_RESET_FINISH || MODE == MODE_RESTORE_RESET || MODE == MODE_STEP || MODE == MODE_STEP_BATCH) && tid == 0) {
        write_request(P.req, b, r);
        if (P.req.stats) {
            if (r.agent >= 0) P.req.stats[3 * b] += 1.0;
            if (MODE == MODE_STEP || MODE == MODE_STEP_BATCH) P.req.stats[3 * b + 1] += r.now - now_before;
            if (MODE == MODE_RESTORE_RESET || MODE == MODE_RESET_FINISH) P.req.stats[3 * b + 2] += 1.0;
        }
    }
}


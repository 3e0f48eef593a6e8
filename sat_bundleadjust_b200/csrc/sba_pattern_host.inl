// Host side of the pattern-major engine (included by sba_ba.cu inside namespace sba, after the shared helpers):
// problem set-up in the internal track order, kernel launches, and the trust-region iteration built on the four fused
// passes of sba_pattern.cuh.  The iteration is the same algorithm as solve_on_device (scipy's trf_no_bounds with an exact
// Gauss-Newton step); what changes is how the work is cut:
//     K2 jvp1 -> damping | K3 eliminate + Schur -> reduce -> Cholesky | K4 back-substitution + Gram scalars -> 2-D step |
//     K1 trial point: cost AND the blocks of the next iteration (an accepted step costs nothing more)
// i.e. 6 launches and 4 Jacobian evaluations per observation and iteration, nothing per-observation stored in HBM.

// shape: 0 = light (jvp1, backsub), 1 = wide (assemble), 2 = narrow (schur)
static PatView pat_view(const sba_problem* p, int shape, int slot = -1)
{
    PatView A;
    A.units = (const PUnit*)p->pt_units[shape];
    A.warp_unit0 = p->pt_warp_unit0[shape];
    A.pts2d = (const double2*)p->pts2d; A.w = p->w; A.cam_static = p->cam_static; A.rpc_tab = p->rpc_tab;
    A.M = p->M; A.P = p->P; A.n_cam_fix = p->n_cam_fix; A.n_cta = p->pt_n_cta;
    static const bool skip = getenv("SBA_PT_SKIP") != nullptr;
    A.debug_skip = skip ? 1 : 0;
    A.cycles = p->pt_cycles ? p->pt_cycles + (size_t)(slot < 0 ? shape : slot) * PT_CTAS * 32 : nullptr;
    return A;
}

constexpr int PT_RED_DOUBLES = 256;     // scratch of cta_reduce_sum: <= 8 values x 32 warps

static size_t pt_smem_common(const sba_problem* p)
{
    return (size_t)p->M * CAMREC_STRIDE + (p->model == MODEL_RPC ? (size_t)p->M * RPC_TAB_STRIDE : 0);
}
static size_t pt_smem_assemble(const sba_problem* p)
{
    const int nv = p->nc * (p->nc + 1) / 2 + p->nc, nw = PT_THREADS / 32;
    return (pt_smem_common(p) + (size_t)nw * p->M * nv + nw * 9 * 33 + PT_RED_DOUBLES) * sizeof(double);
}
static size_t pt_smem_jvp1(const sba_problem* p)
{
    return (pt_smem_common(p) + (size_t)p->M * p->nc + PT_RED_DOUBLES) * sizeof(double);
}
static size_t pt_smem_schur(const sba_problem* p)
{
    return (pt_smem_common(p) + (size_t)(PT_THREADS_SCHUR / 32) * (32 * (p->nc * 3 + 1) + PT_RC * 3)) * sizeof(double);
}
static size_t pt_smem_backsub(const sba_problem* p)
{
    return (pt_smem_common(p) + 2 * (size_t)p->M * p->nc + (PT_THREADS_LIGHT / 32) * 3 * 32 + PT_RED_DOUBLES) * sizeof(double);
}

#define PT_DISPATCH(p, MACRO)                                                           \
    switch ((p)->model * 16 + (p)->nc) {                                                \
    case MODEL_PERSPECTIVE * 16 + 3: MACRO(MODEL_PERSPECTIVE, 3); break;                \
    case MODEL_PERSPECTIVE * 16 + 6: MACRO(MODEL_PERSPECTIVE, 6); break;                \
    case MODEL_AFFINE * 16 + 3: MACRO(MODEL_AFFINE, 3); break;                          \
    case MODEL_AFFINE * 16 + 5: MACRO(MODEL_AFFINE, 5); break;                          \
    case MODEL_RPC * 16 + 3: MACRO(MODEL_RPC, 3); break;                                \
    case MODEL_RPC * 16 + 6: MACRO(MODEL_RPC, 6); break;                                \
    default: set_error("pattern engine: unsupported (cam_model, n_params)"); return SBA_E_INVALID; \
    }

// kernels that need more than 48 KB of dynamic shared memory must opt in, once per process and device
static int pt_set_smem_attributes(const sba_problem* p)
{
#define L(MODEL, NC)                                                                                                         \
    SBA_CUDA(cudaFuncSetAttribute(k_pt_assemble<MODEL, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pt_smem_assemble(p))); \
    SBA_CUDA(cudaFuncSetAttribute(k_pt_jvp1<MODEL, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pt_smem_jvp1(p)));   \
    SBA_CUDA(cudaFuncSetAttribute(k_pt_schur<MODEL, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pt_smem_schur(p))); \
    SBA_CUDA(cudaFuncSetAttribute(k_pt_schur_mma<MODEL, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pt_smem_schur(p))); \
    SBA_CUDA(cudaFuncSetAttribute(k_pt_backsub<MODEL, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pt_smem_backsub(p)))
    PT_DISPATCH(p, L);
#undef L
    return SBA_OK;
}

static bool pattern_engine_applicable(const sba_problem* p, bool forced = false)
{
    if (const char* e = getenv("SBA_ENGINE")) if (std::strcmp(e, "generic") == 0 && !forced) return false;
    if (p->n_common != 0 || p->nc > 6 || p->M * p->nc > PT_MAX_NS || p->M > 64) return false;
    return pt_smem_schur(p) <= (size_t)220 * 1024 && pt_smem_assemble(p) <= (size_t)220 * 1024 && pt_smem_backsub(p) <= (size_t)220 * 1024;
}

// ---- boundary: caller's order <-> internal order ------------------------------------------------------------------
// x_ext (device, caller's layout) -> internal vector dst
static int pt_x_in(sba_problem* p, const double* x_ext_dev, double* dst)
{
    k_pt_points_in<<<grid_for(p->n, 256, NUM_SMS * 8), 256, 0, p->stream>>>(x_ext_dev, p->trk_new2old, p->N, p->M * p->nc, dst);
    return check_launch(p);
}
static int pt_x_out(sba_problem* p, const double* src, double* x_ext_dev)
{
    k_pt_points_out<<<grid_for(p->n, 256, NUM_SMS * 8), 256, 0, p->stream>>>(src, p->trk_new2old, p->N, p->M * p->nc, x_ext_dev);
    return check_launch(p);
}
static int pt_obs_out(sba_problem* p, const double* src_int, int width, double* dst_ext_dev)
{
    k_pt_obs_out<<<grid_for(p->K * width, 256, NUM_SMS * 8), 256, 0, p->stream>>>(src_int, p->obs_new2old, p->K, width, dst_ext_dev);
    return check_launch(p);
}

// ---- launches ----------------------------------------------------------------------------------------------------------
// K1 + reduction: trial point from (x, g, delta, step coefficients on the device) into the `new` buffer set
static int pt_run_assemble(sba_problem* p, int initial, int first, int loss, double f_scale)
{
    const int ns = p->M * p->nc;
    const size_t cs_count = (size_t)ns * p->nc + ns + 1;
    const CommFused cf = comm_fused(p, (long long)cs_count);
#define L(MODEL, NC)                                                                                                          \
    k_pt_assemble<MODEL, NC><<<p->pt_n_cta, PT_THREADS, pt_smem_assemble(p), p->stream>>>(                                     \
        pat_view(p, 1), p->x, p->g, p->idsq, p->idsqc, p->delta, p->scal, initial, ns, loss, f_scale, p->x_new, p->camrec_new,    \
        p->V2, p->g2, (double2*)p->osc2, p->pt_partials);                                                                      \
    SBA_TRY(check_launch(p));                                                                                                  \
    k_pt_reduce_assemble<NC><<<(p->M * (NC * (NC + 1) / 2 + NC) + 1 + 7) / 8, 256, 0, p->stream>>>(                          \
        p->pt_partials, p->pt_n_cta, p->M, p->camsys2, p->world == 1 || cf.on, p->dsqc, first, p->dsqc2, p->idsqc2, p->g2, p->scal, \
        p->counters + 4, cf)
    PT_DISPATCH(p, L);
#undef L
    SBA_TRY(check_launch(p));
    if (p->world > 1 && !cf.on) {
        SBA_TRY(allreduce_any(p, p->camsys2, (long long)cs_count));
        k_pt_cam_scale<<<1, 256, 0, p->stream>>>(p->camsys2, p->dsqc, first, p->M, p->nc, p->dsqc2, p->idsqc2, p->g2, p->scal);
        SBA_TRY(check_launch(p));
    }
    return SBA_OK;
}

// the trial point becomes the current point
static void pt_accept(sba_problem* p)
{
    std::swap(p->x, p->x_new); std::swap(p->camrec, p->camrec_new);
    std::swap(p->V, p->V2); std::swap(p->g, p->g2); std::swap(p->camsys, p->camsys2);
    std::swap(p->dsqc, p->dsqc2); std::swap(p->idsqc, p->idsqc2); std::swap(p->osc, p->osc2);
}

static int pt_run_jvp1(sba_problem* p, int first, int loss, double f_scale, double delta_arg)
{
    const int ns = p->M * p->nc;
    SBA_CUDA(cudaMemsetAsync(p->scal + SC_GMAX_SLOTS, 0, 16 * sizeof(double), p->stream));
    const CommFused cf = comm_fused(p, SC_GGN - SC_COST);
    const int fold = p->world == 1 || cf.on;
#define L(MODEL, NC)                                                                                                       \
    k_pt_jvp1<MODEL, NC><<<p->pt_n_cta, PT_THREADS_LIGHT, pt_smem_jvp1(p), p->stream>>>(                                    \
        pat_view(p, 0), p->x, p->camrec, p->V, p->g, p->dsqc, p->idsqc, p->dsq, p->idsq, (const double2*)p->osc, first, ns,   \
        p->rank == 0, p->rank, p->red_partials, p->counters + 2, p->scal, fold, delta_arg, cf)
    PT_DISPATCH(p, L);
#undef L
    SBA_TRY(check_launch(p));
    if (!fold) {
        SBA_TRY(allreduce_scal(p, SC_COST, SC_GGN - SC_COST));
        k_pt_control_reg<<<1, 32, 0, p->stream>>>(p->scal, delta_arg, -1.0);
        SBA_TRY(check_launch(p));
    }
    return SBA_OK;
}

// K3 + reduction: the reduced camera system for the damping in scal[SC_REG]
static int pt_run_schur(sba_problem* p, int loss, double f_scale)
{
    const int ns = p->M * p->nc, nS = p->nc * p->nc * (p->M * (p->M + 1) / 2);
    SBA_CUDA(cudaMemsetAsync(p->scal + SC_BAD_POINTS, 0, 2 * sizeof(double), p->stream));
    const CommFused cf = comm_fused(p, (long long)nS + ns);
#define L(MODEL, NC)                                                                                                        \
    if (p->pt_schur_mma)                                                                                                     \
        k_pt_schur_mma<MODEL, NC><<<p->pt_n_cta, PT_THREADS_SCHUR, pt_smem_schur(p), p->stream>>>(                           \
            pat_view(p, 2), p->x, p->camrec, p->V, p->g, p->dsq, (const double2*)p->osc, p->scal, ns, p->pt_records, p->pt_partials, \
            p->scal + SC_BAD_POINTS);                                                                                        \
    else                                                                                                                     \
        k_pt_schur<MODEL, NC><<<p->pt_n_cta, PT_THREADS_SCHUR, pt_smem_schur(p), p->stream>>>(                               \
            pat_view(p, 2), p->x, p->camrec, p->V, p->g, p->dsq, (const double2*)p->osc, p->scal, ns, p->pt_records, p->pt_partials, \
            p->scal + SC_BAD_POINTS);                                                                                        \
    SBA_TRY(check_launch(p));                                                                                                \
    k_pt_reduce_schur<NC><<<std::min(PT_CTAS, (nS + ns + 15) / 16), 512, 0, p->stream>>>(                                    \
        p->pt_partials, p->pt_n_cta, p->M, p->n_cam_fix, p->camsys, p->dsqc, p->scal, p->rank == 0, p->S, cf, p->counters + 6)
    PT_DISPATCH(p, L);
#undef L
    SBA_TRY(check_launch(p));
    if (!cf.on) SBA_TRY(allreduce_any(p, p->S, (long long)ns * ns + ns));
    return SBA_OK;
}

static int pt_run_backsub(sba_problem* p, int loss, double f_scale)
{
    const int ns = p->M * p->nc;
    const CommFused cf = comm_fused(p, 7);
    const int fold = p->world == 1 || cf.on;
#define L(MODEL, NC)                                                                                                      \
    k_pt_backsub<MODEL, NC><<<p->pt_n_cta, PT_THREADS_LIGHT, pt_smem_backsub(p), p->stream>>>(                             \
        pat_view(p, 0, 3), p->x, p->camrec, p->V, p->g, p->dsq, p->idsq, p->dsqc, p->idsqc, (const double2*)p->osc, p->delta, ns, \
        p->rank == 0, p->red_partials, p->counters + 3, p->scal, fold, cf)
    PT_DISPATCH(p, L);
#undef L
    SBA_TRY(check_launch(p));
    if (!fold) {
        SBA_TRY(allreduce_scal(p, SC_P_GD, 7));
        k_pt_control_tr2d<<<1, 32, 0, p->stream>>>(p->scal, -1.0);
        SBA_TRY(check_launch(p));
    }
    return SBA_OK;
}

// state of a fresh solve: scales unset, no step, both x buffers hold the start vector (tracks without observations are
// never written by the kernels)
static int pt_reset_state(sba_problem* p)
{
    const size_t n = (size_t)p->n;
    SBA_CUDA(cudaMemsetAsync(p->scal, 0, SC_COUNT * sizeof(double), p->stream));
    SBA_CUDA(cudaMemcpyAsync(p->x_new, p->x, n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    SBA_CUDA(cudaMemsetAsync(p->delta, 0, n * sizeof(double), p->stream));
    return SBA_OK;
}

// SBA_PT_CYCLES=1: load balance of the static assignment -- clocks every warp spent in its unit loop during the LAST launch of each kernel
static void pt_print_cycles(sba_problem* p)
{
    if (!p->pt_cycles) return;
    std::vector<long long> h((size_t)5 * PT_CTAS * 32);
    if (cudaMemcpy(h.data(), p->pt_cycles, h.size() * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    const char* names[4] = {"K2 jvp1 (16 warps)", "K1 assemble (16 warps)", "K3 schur (12 warps)", "K4 backsub (16 warps)"};
    const int warps[4] = {PT_THREADS_LIGHT / 32, PT_THREADS / 32, PT_THREADS_SCHUR / 32, PT_THREADS_LIGHT / 32};
    if (const char* path = getenv("SBA_PT_CYCLES_FILE")) {      // raw per-warp clocks: 4 kernels x 148 CTAs x 32 slots + the SM id of every CTA
        if (FILE* f = fopen(path, "wb")) { fwrite(h.data(), sizeof(long long), h.size(), f); fclose(f); }
    }
    if (getenv("SBA_PT_CYCLES_DUMP")) {
        for (int k = 0; k < 4; ++k) {
            fprintf(stderr, "[sba cycles dump] k%d:", k);
            for (int c = 0; c < p->pt_n_cta; ++c) { double a = 0; for (int w = 0; w < warps[k]; ++w) a += (double)h[(size_t)k * PT_CTAS * 32 + (size_t)c * warps[k] + w]; fprintf(stderr, " %d:%.0f", (int)h[(size_t)4 * PT_CTAS * 32 + c], a / warps[k] / 100.0); }
            fprintf(stderr, "\n");
        }
    }
    for (int k = 0; k < 4; ++k) {
        double wsum = 0, wmax = 0, csum = 0, cmax = 0, cmin = 1e30;
        for (int c = 0; c < p->pt_n_cta; ++c) {
            double cta = 0;
            for (int w = 0; w < warps[k]; ++w) {
                const double v = (double)h[(size_t)k * PT_CTAS * 32 + (size_t)c * warps[k] + w];
                wsum += v; wmax = std::max(wmax, v); cta = std::max(cta, v);
            }
            csum += cta; cmax = std::max(cmax, cta); cmin = std::min(cmin, cta);
        }
        const int nwarp = p->pt_n_cta * warps[k];
        fprintf(stderr, "[sba cycles] %-24s per warp: avg %.0f max %.0f (max/avg %.2f) | per CTA (slowest warp): min %.0f avg %.0f max %.0f (max/avg %.2f)\n",
                names[k], wsum / nwarp, wmax, wmax / (wsum / nwarp), cmin, csum / p->pt_n_cta, cmax, cmax / (csum / p->pt_n_cta));
    }
}

// ---- the iteration -------------------------------------------------------------------------------------------------------
static int solve_pattern(sba_problem* p, const sba_solve_opts* o, sba_solve_info* info)
{
    const int loss = o->loss;
    const double fs = o->f_scale;
    std::memset(info, 0, sizeof(*info));
    if (o->max_nfev < 1) { set_error("max_nfev must be >= 1"); return SBA_E_INVALID; }
    p->launches = 0;
    SBA_TRY(pt_reset_state(p));
    SBA_CUDA(cudaEventRecord(p->ev0, p->stream));
    // residual, cost and blocks at x0 (the same kernel evaluates every later trial point)
    SBA_TRY(pt_run_assemble(p, 1, 1, loss, fs));
    SBA_TRY(fetch_scal(p));
    double cost = p->h_scal[SC_COST_NEW];
    if (!std::isfinite(cost)) { set_error("Residuals are not finite in the initial point."); return SBA_E_NUMERIC; }
    info->cost_init = cost;
    pt_accept(p);
    int nfev = 1, njev = 1, iteration = 0, status = -1, chol_retries = 0;
    double Delta = -1.0, g_norm = 0.0;
    bool first = true;
    PhaseTimer tm, it_tm;
    tm.p = it_tm.p = p;
    p->ev_used = 0;
    if (o->l2_flush_bytes > 0 && (size_t)o->l2_flush_bytes > p->flush_bytes) {
        if (p->flush_buf) cudaFree(p->flush_buf);
        p->flush_buf = nullptr; p->flush_bytes = 0;
        SBA_CUDA(cudaMalloc(&p->flush_buf, (size_t)o->l2_flush_bytes));
        p->flush_bytes = (size_t)o->l2_flush_bytes;
    }
    const int ns = p->M * p->nc;
    // trial point for radius `radius` (< 0: the step already on the device): step coefficients, then K1
    auto enqueue_trial = [&](double radius) -> int {
        tm.begin(SBA_PH_STEP_EVAL);
        if (radius >= 0.0) {
            k_pt_control_tr2d<<<1, 32, 0, p->stream>>>(p->scal, radius);
            SBA_TRY(check_launch(p));
        }
        SBA_TRY(pt_run_assemble(p, 0, 0, loss, fs));
        tm.end();
        return SBA_OK;
    };
    auto enqueue_iteration = [&](double reg_override) -> int {
        if (reg_override < 0.0) {
            tm.begin(SBA_PH_SCALE_JVP);
            SBA_TRY(pt_run_jvp1(p, first ? 1 : 0, loss, fs, Delta));
            tm.end();
        } else {
            k_pt_control_reg<<<1, 32, 0, p->stream>>>(p->scal, Delta, reg_override);
            SBA_TRY(check_launch(p));
        }
        tm.begin(SBA_PH_SCHUR);
        SBA_TRY(pt_run_schur(p, loss, fs));
        tm.end();
        tm.begin(SBA_PH_CHOLESKY);
        SBA_TRY(launch_cholesky_solve(p->S, p->S + (size_t)ns * ns, p->delta, ns, p->scal + SC_CHOL_FAIL, p->chol_work, p->stream, false));
        p->launches++;
        tm.end();
        tm.begin(SBA_PH_BACKSUB);
        SBA_TRY(pt_run_backsub(p, loss, fs));
        tm.end();
        return enqueue_trial(-1.0);
    };

    while (true) {
        if (o->max_iterations > 0 && iteration >= o->max_iterations) break;
        if (nfev >= o->max_nfev) break;
        it_tm.on = iteration >= o->timed_from && (o->timed_from > 0 || o->l2_flush_bytes > 0 || o->max_iterations > 0);
        tm.on = it_tm.on && !o->no_phase_timing;
        if (o->l2_flush_bytes > 0)
            SBA_CUDA(cudaMemsetAsync(p->flush_buf, iteration & 0xff, (size_t)o->l2_flush_bytes, p->stream));
        it_tm.begin(-1);
        SBA_TRY(enqueue_iteration(-1.0));
        SBA_TRY(fetch_scal(p));
        const double* h = p->h_scal;
        if (h[SC_COMM_FAIL] != 0.0) { set_error("peer-memory all-reduce timed out (a rank is missing)"); return SBA_E_CUDA; }
        for (int attempt = 0; h[SC_CHOL_FAIL] != 0.0 || !std::isfinite(h[SC_P_GD]) || !std::isfinite(h[SC_P_C22]); ++attempt) {
            if (attempt >= 30) { set_error("reduced camera system could not be factorised"); return SBA_E_NUMERIC; }
            ++chol_retries;
            SBA_TRY(enqueue_iteration(std::max(h[SC_REG] * 10.0, 1e-12)));
            SBA_TRY(fetch_scal(p));
            h = p->h_scal;
        }
        first = false;
        Delta = h[SC_DELTA];
        const double x_norm = std::sqrt(h[SC_XX]);
        g_norm = 0.0;
        for (int r = 0; r < 16; ++r) g_norm = std::max(g_norm, h[SC_GMAX_SLOTS + r]);
        if (o->verbose >= 2)
            printf("[sba] it %3d nfev %3d cost %.10e |g|inf %.3e Delta %.3e reg %.3e\n", iteration, nfev, cost, g_norm, Delta, h[SC_REG]);
        if (g_norm < o->gtol) { status = 1; break; }

        double actual_reduction = -1.0, cost_new = cost;
        int term = -1;
        while (true) {
            ++nfev;
            cost_new = h[SC_COST_NEW];
            const double predicted = h[SC_PRED], step_h_norm = h[SC_STEPH], step_norm = h[SC_STEPN];
            if (!std::isfinite(cost_new)) {
                Delta = 0.25 * step_h_norm;
            } else {
                actual_reduction = cost - cost_new;
                double ratio;
                if (predicted > 0.0) ratio = actual_reduction / predicted;
                else if (predicted == 0.0 && actual_reduction == 0.0) ratio = 1.0;
                else ratio = 0.0;
                double Delta_new = Delta;
                if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
                else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = 2.0 * Delta;
                const bool ftol_ok = actual_reduction < o->ftol * cost && ratio > 0.25;
                const bool xtol_ok = step_norm < o->xtol * (o->xtol + x_norm);
                if (ftol_ok && xtol_ok) term = 4; else if (ftol_ok) term = 2; else if (xtol_ok) term = 3;
                if (term >= 0) break;
                Delta = Delta_new;
            }
            if (actual_reduction > 0.0 || nfev >= o->max_nfev) break;
            SBA_TRY(enqueue_trial(Delta));
            SBA_TRY(fetch_scal(p));
            h = p->h_scal;
        }
        if (actual_reduction > 0.0) {
            pt_accept(p);          // the trial evaluation already holds the blocks of the new point
            cost = cost_new;
            ++njev;
        }
        it_tm.end();
        ++iteration;
        if (term >= 0) { status = term; break; }
    }
    if (status < 0) status = 0;
    SBA_CUDA(cudaEventRecord(p->ev1, p->stream));
    SBA_CUDA(cudaStreamSynchronize(p->stream));
    float ms = 0.f;
    SBA_CUDA(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    info->status = status; info->nfev = nfev; info->njev = njev; info->iterations = iteration;
    info->cost = cost; info->optimality = g_norm; info->solve_ms = ms; info->chol_retries = chol_retries;
    info->gpu_launches = p->launches;
    info->explicit_subspace_passes = iteration;
    tm.resolve(info);
    it_tm.resolve(info);
    return SBA_OK;
}

// ---- problem set-up ------------------------------------------------------------------------------------------------------
// hidx: validated int32 indices of the caller (build_host_index); lay: the pattern layout built from them
static int pattern_create(sba_problem* p, const sba_problem_desc* d, const HostIndex& hidx, const PatternLayout& lay)
{
    const int M = p->M, N = p->N, nc = p->nc;
    const int64_t K = p->K;
    cudaStream_t s = p->stream;
    p->engine = 1;
    p->use_pcg = false;
    p->pt_n_cta = lay.n_cta;
    p->n_pts_fix_int = lay.n_frozen_tracks;
    p->h_trk_new2old = lay.trk_new2old;
    // internal-order indices for the generic per-observation kernels (residual output, Jacobian blocks) and the observation
    // permutation are derived on the device from the track permutation
    int *d_cam_ext = nullptr, *d_tp_old = nullptr;
    SBA_TRY(dev_upload(p, &d_cam_ext, hidx.cam, s));
    SBA_TRY(dev_upload(p, &d_tp_old, hidx.track_ptr, s));
    SBA_TRY(dev_alloc(p, &p->cam_ind, (size_t)K)); SBA_TRY(dev_alloc(p, &p->pts_ind, (size_t)K));
    SBA_TRY(dev_alloc(p, &p->obs_new2old, (size_t)K));
    SBA_TRY(dev_upload(p, &p->track_ptr, lay.track_ptr, s));
    SBA_TRY(dev_upload(p, &p->trk_new2old, lay.trk_new2old, s));
    k_pt_build_obs<<<grid_for(N, 256, NUM_SMS * 8), 256, 0, s>>>(p->trk_new2old, d_tp_old, p->track_ptr, d_cam_ext, N, p->obs_new2old,
                                                                p->cam_ind, p->pts_ind);
    SBA_CUDA(cudaGetLastError());
    for (int k = 0; k < 3; ++k) {
        const PatternAssignment& as = k == 0 ? lay.light : (k == 1 ? lay.wide : lay.narrow);
        SBA_TRY(dev_upload(p, &p->pt_warp_unit0[k], as.warp_unit0, s));
        PUnit* du = nullptr;
        SBA_TRY(dev_alloc(p, &du, as.units.size()));
        SBA_CUDA(cudaMemcpyAsync(du, as.units.data(), as.units.size() * sizeof(PUnit), cudaMemcpyHostToDevice, s));
        p->pt_units[k] = du;
    }
    // warp tiles of the generic per-track kernels are not used by this engine
    p->n_tiles = 0;
    if (!p->pt_obs_uploaded) {
        SBA_TRY(dev_alloc(p, &p->r_out, 2 * (size_t)K));
        SBA_TRY(dev_alloc(p, &p->err_out, (size_t)K));
    }
    SBA_TRY(dev_alloc(p, &p->osc, 2 * (size_t)K)); SBA_TRY(dev_alloc(p, &p->osc2, 2 * (size_t)K));
    SBA_TRY(dev_alloc(p, &p->r_int, 2 * (size_t)K));
    SBA_TRY(dev_alloc(p, &p->e_int, (size_t)K));
    // observations and weights: uploaded in the caller's order, gathered into the internal order on the device
    SBA_TRY(dev_alloc(p, &p->pts2d, 2 * (size_t)K));
    SBA_TRY(dev_alloc(p, &p->w, (size_t)K));
    if (!p->pt_obs_uploaded) {      // else: already on their way (issued by problem_create_impl's helper thread on this stream)
        SBA_CUDA(cudaMemcpyAsync(p->r_out, d->pts2d, 2 * (size_t)K * sizeof(double), cudaMemcpyHostToDevice, s));
        SBA_CUDA(cudaMemcpyAsync(p->err_out, d->pts2d_w, (size_t)K * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    k_pt_obs_in<<<grid_for(2 * K, 256, NUM_SMS * 8), 256, 0, s>>>(p->r_out, p->obs_new2old, K, 2, p->pts2d);
    SBA_CUDA(cudaGetLastError());
    k_pt_obs_in<<<grid_for(K, 256, NUM_SMS * 8), 256, 0, s>>>(p->err_out, p->obs_new2old, K, 1, p->w);
    SBA_CUDA(cudaGetLastError());
    SBA_TRY(dev_alloc(p, &p->cam_static, (size_t)M * p->P));
    SBA_CUDA(cudaMemcpyAsync(p->cam_static, d->cam_params, (size_t)M * p->P * sizeof(double), cudaMemcpyHostToDevice, s));
    if (p->model == MODEL_RPC) {
        SBA_TRY(dev_alloc(p, &p->rpc_tab, (size_t)M * RPC_TAB_STRIDE));
        SBA_CUDA(cudaMemcpyAsync(p->rpc_tab, d->rpc_coefs, (size_t)M * RPC_TAB_STRIDE * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    // iteration state
    const size_t n = (size_t)p->n, ns = (size_t)M * nc;
    SBA_TRY(dev_alloc(p, &p->x, n)); SBA_TRY(dev_alloc(p, &p->x_new, n)); SBA_TRY(dev_alloc(p, &p->io_x, n));
    SBA_TRY(dev_alloc(p, &p->g, n)); SBA_TRY(dev_alloc(p, &p->g2, n));
    SBA_TRY(dev_alloc(p, &p->dsq, n)); SBA_TRY(dev_alloc(p, &p->idsq, n)); SBA_TRY(dev_alloc(p, &p->delta, n));
    SBA_TRY(dev_alloc(p, &p->V, 6 * (size_t)N)); SBA_TRY(dev_alloc(p, &p->V2, 6 * (size_t)N));
    SBA_TRY(dev_alloc(p, &p->camrec, (size_t)M * CAMREC_STRIDE)); SBA_TRY(dev_alloc(p, &p->camrec_new, (size_t)M * CAMREC_STRIDE));
    for (double* buf : {p->g, p->g2, p->delta, p->x_new}) SBA_CUDA(cudaMemsetAsync(buf, 0, n * sizeof(double), s));
    SBA_CUDA(cudaMemsetAsync(p->V, 0, 6 * (size_t)N * sizeof(double), s));
    SBA_CUDA(cudaMemsetAsync(p->V2, 0, 6 * (size_t)N * sizeof(double), s));
    k_pt_fill<<<grid_for((long long)n, 256, NUM_SMS * 8), 256, 0, s>>>(p->dsq, 1.0, (long long)n);
    k_pt_fill<<<grid_for((long long)n, 256, NUM_SMS * 8), 256, 0, s>>>(p->idsq, 1.0, (long long)n);
    SBA_CUDA(cudaGetLastError());
    const size_t cs = ns * nc + ns + 1;
    SBA_TRY(dev_alloc(p, &p->camsys, cs)); SBA_TRY(dev_alloc(p, &p->camsys2, cs));
    p->camsys_local = p->camsys;
    SBA_TRY(dev_alloc(p, &p->dsqc, ns)); SBA_TRY(dev_alloc(p, &p->dsqc2, ns));
    SBA_TRY(dev_alloc(p, &p->idsqc, ns)); SBA_TRY(dev_alloc(p, &p->idsqc2, ns));
    for (double* buf : {p->dsqc, p->dsqc2, p->idsqc, p->idsqc2}) {
        k_pt_fill<<<1, 256, 0, s>>>(buf, 1.0, (long long)ns);
        SBA_CUDA(cudaGetLastError());
    }
    SBA_TRY(dev_alloc(p, &p->S, ns * ns + ns));
    SBA_TRY(dev_alloc(p, &p->chol_work, (size_t)34 * (ns + 32)));
    const size_t nv = (size_t)nc * (nc + 1) / 2 + nc, nS = (size_t)nc * nc * ((size_t)M * (M + 1) / 2);
    SBA_TRY(dev_alloc(p, &p->pt_partials, std::max((size_t)M * nv + 1, nS + ns) * lay.n_cta));
    if (getenv("SBA_PT_CYCLES")) {
        SBA_TRY(dev_alloc(p, &p->pt_cycles, (size_t)5 * PT_CTAS * 32));
        SBA_CUDA(cudaMemsetAsync(p->pt_cycles, 0, (size_t)5 * PT_CTAS * 32 * sizeof(long long), s));
    }
    SBA_TRY(dev_alloc(p, &p->pt_records, (size_t)std::max(1, lay.narrow.n_records) *
                                             (p->pt_schur_mma ? (size_t)PT_MMA_TILES * 64 + nc * 32 : (size_t)(2 * PT_RC * nc + nc) * 32)));
    SBA_TRY(dev_alloc(p, &p->red_partials, (size_t)std::max(NUM_SMS * 16, lay.n_cta + 1) * 8));
    SBA_TRY(dev_alloc(p, &p->counters, 16));
    SBA_CUDA(cudaMemsetAsync(p->counters, 0, 16 * sizeof(unsigned), s));
    SBA_TRY(dev_alloc(p, &p->scal, SC_COUNT));
    SBA_CUDA(cudaMemsetAsync(p->scal, 0, SC_COUNT * sizeof(double), s));
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (!g_pinned_pool.empty()) { p->h_scal = g_pinned_pool.back(); g_pinned_pool.pop_back(); }
    }
    if (!p->h_scal) SBA_CUDA(cudaMallocHost((void**)&p->h_scal, H_SCAL_COUNT * sizeof(double)));
    SBA_CUDA(cudaEventCreate(&p->ev0));
    SBA_CUDA(cudaEventCreate(&p->ev1));
    SBA_TRY(pt_set_smem_attributes(p));
    SBA_CUDA(cudaStreamSynchronize(s));   // host staging vectors go out of scope
    return SBA_OK;
}

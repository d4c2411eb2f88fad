// cp_host.inl -- host control loop of the constrained solver (included by cv_cp.cu).
//
// Mirrors CPSolver::{new, solve, solve_r, backtrack} (reference src/viterbi_solver/cp.rs:20-30,
// 85-143): the recursion over components and states, the pruning test `ub > best_obj` and the
// explored-node counter run here on the host exactly as in the reference; every array operation
// (sweeps, fix-ups, bound terms, their ordered sum, the backtrack) is a kernel of cp_kernels.cuh.
// Nothing here computes Viterbi values on the CPU.

static int g_sum_parallel_min = 4096;
static int g_sum_force = -1;   // CV_CP_SUM=0|1|2 forces the plain loop / block-structured / single-CTA exact sum
static double g_cp_prof[6];   // CV_CP_PROF=1: host-timed phases of a node (sweep, fixup, terms, sum, readback)

// The exchange buffer of one rank (cp_dist.cuh): [flags][maps x2][terms x2][solution], mapped into every rank of
// the group through CUDA IPC.  Created once per (model, group, capacity) and reused across solves.
struct cv_cp_dist {
    cv_hmm *h = nullptr;
    int rank = 0, R = 1, mapw = 0;
    int64_t cap_N = 0, cap_terms = 0;
    size_t maps_off = 0, terms_off = 0, sol_off = 0, bytes = 0;
    void *mine = nullptr;
    PeerTab tab{};
    bool connected = false;
    unsigned int epoch = 0;                      // exchanges so far; identical on every rank
};

namespace {

struct CpRun {
    cv_hmm *h;
    CpParams p;
    cudaStream_t st;
    int32_t ncomp;
    uint64_t max_nodes, explored = 0, steps = 0;
    double best_obj = -std::numeric_limits<double>::infinity();
    std::vector<int64_t> cons_off;               // positions of component c: cons_pos[cons_off[c] .. cons_off[c+1])
    int64_t *d_cons_pos = nullptr;               // all components back to back = term order of cp.rs:104-116
    int32_t *d_term_comp = nullptr;
    std::vector<int64_t> seg_off;                // sweep segments of depth c: [seg_off[c], seg_off[c+1])
    std::vector<uint64_t> seg_steps;             // rows swept per node of depth c
    int64_t *d_seg_from = nullptr; int32_t *d_seg_len = nullptr;
    double *d_terms = nullptr, *d_ub = nullptr; int *d_end = nullptr; unsigned int *d_counter = nullptr;
    SumWs sum_ws;                                // block statistics / functions of the exact-order sum
    // sharding (cp_dist.cuh): this rank's cut pair; R == 1 => lo = 0, hi = N and nothing is exchanged
    cv_cp_dist *dx = nullptr;
    int rank = 0, R = 1;
    int64_t lo = 0, hi = 0;
    std::vector<int64_t> fix_off;                // fix-up positions of depth c: d_fix_pos[fix_off[c] .. fix_off[c+1])
    int64_t *d_fix_pos = nullptr;
    std::vector<int64_t> lterm_off;              // local bound terms of components 0..c: first lterm_off[c+1] entries
    int64_t *d_lterm_pos = nullptr; int32_t *d_lterm_comp = nullptr, *d_lterm_gidx = nullptr;
    unsigned int *d_done = nullptr; int *d_err = nullptr;
    uint64_t n_term_x = 0, n_map_x = 0;          // exchanges of each kind so far in this solve (buffer parity)
    unsigned long long node_seq = 0;             // sequence number of the node whose bound the host polls for
    int sweep_blocks_per_sm = 8;                 // co-resident CTAs of the sweep kernel (occupancy query)
    bool poll = true;                            // CV_CP_HOSTPOLL=0: copy + stream synchronise instead
    psi_t *d_F = nullptr; int *d_entry = nullptr; uint64_t *d_sol = nullptr;
    // leaf batch (cp_leaf_group): the siblings of the last component evaluated together
    bool leaf_batch = false;
    int nleaf = 0;                               // positions of the last component (= its sweeps)
    double *d_leaf_last = nullptr;               // [K siblings][nleaf][K] last row of every sibling's sweeps
    psi_t *d_c2 = nullptr;                       // [K siblings][nleaf] C2 backpointers of every sibling
    double *d_leaf_terms = nullptr, *d_leaf_ub = nullptr;   // [K][leaf_term_stride], [K]
    long long leaf_term_stride = 0;
    int32_t *d_prev_seg = nullptr, *d_leaf_idx = nullptr; int64_t *d_leaf_pos = nullptr;
    SumWs leaf_ws;
    double *h_leaf_ub = nullptr;                 // pinned, [K]
    double *h_ub = nullptr;                      // pinned
    int err = CV_OK;
    size_t smem;
};

// ub = ((0.0 + x_0) + x_1) + ... in order (cp.rs:103-116): plain loop for short lists, the block-structured
// exact-order kernels for long ones, the single-CTA binade scan beyond SUM_MAX_BLOCKS blocks.
// nlists > 1: a batch of lists (leaf batch), list y at terms + y * term_stride with its block workspace at
// + y * ws.blk_stride, result y to ub.dev[y] (ub.host must be null); every kernel runs with gridDim.y = nlists.
int cp_launch_sum(const double *terms, int nterms, const UbSink &ub, unsigned int *counter, const SumWs &ws, bool have_stats,
                  int force, cudaStream_t st, const PeerWait &pw, int nlists = 1, long long term_stride = 0)
{
    const int nblk = (nterms + SUM_BLK - 1) / SUM_BLK;
    int kind = nterms < g_sum_parallel_min ? 0 : (nblk > SUM_MAX_BLOCKS ? 2 : 1);
    if (force >= 0) kind = (force == 1 && nblk > SUM_MAX_BLOCKS) ? 2 : force;
    if (kind == 1 && nblk == 0) kind = 0;
    const SumBatch batch{nlists > 1 ? term_stride : 0, nlists > 1 ? ws.blk_stride : 0};
    const unsigned ny = (unsigned)std::max(nlists, 1);
    if (kind == 0) { cp_sum_kernel<<<dim3(1, ny), 256, 0, st>>>(terms, nterms, ub, counter, pw, batch); g_launches++; }
    else if (kind == 2) {
        if (pw.R > 1) { cp_peer_wait_kernel<<<1, 32, 0, st>>>(pw); g_launches++; }
        cp_sum_exact_kernel<<<dim3(1, ny), QS_THREADS, 0, st>>>(terms, nterms, ub, counter, batch);
        g_launches++;
    } else {
        if (!have_stats || pw.R > 1) { cp_sum_stats_kernel<<<dim3(nblk, ny), SUM_BLK, 0, st>>>(terms, nterms, ws.bsum, ws.bflag, pw, batch); g_launches++; }
        const double *bpre = nullptr;
        if (nblk >= SUM_PREFIX_MIN) { cp_sum_prefix_kernel<<<dim3(1, ny), 1024, 0, st>>>(ws.bsum, nblk, ws.bpre, batch); g_launches++; bpre = ws.bpre; }
        cp_sum_blockfn_kernel<<<dim3(nblk, ny), SUM_BLK, 0, st>>>(terms, nterms, ws.bsum, bpre, ws.bexp, ws.bfn, batch);
        cp_sum_chain_kernel<<<dim3(1, ny), SUMC_THREADS, sum_chain_smem_bytes(nblk), st>>>(terms, nterms, nblk, ws.bflag, ws.bexp, ws.bfn, ub, counter, batch);
        g_launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    return CV_OK;
}

// leaf_nsib > 0: leaf batch -- the sweeps of siblings 0 .. leaf_nsib-1 in one launch, last rows only (cp_kernels.cuh)
int cp_sweep(CpRun &r, int64_t seg_begin, int64_t nseg, int node, int init_mode, int leaf_nsib = 0)
{
    if (nseg <= 0) return CV_OK;
    CpSweepArgs a;
    a.seg_from = r.d_seg_from + seg_begin; a.seg_len = r.d_seg_len + seg_begin;
    a.nseg = (int)nseg; a.ntiles = (int)((nseg + 63) / 64); a.node = node; a.init_mode = init_mode;
    a.tile_counter = r.d_counter;
    a.leaf_nsib = leaf_nsib; a.leaf_last = r.d_leaf_last;
    const int64_t ntask = nseg * std::max(leaf_nsib, 1);
    // (the segment counter is zeroed by the previous node's sum kernel / the set-up memset)
    if (r.p.K > SMALL_K_MAX) {                                  // generic path: states looped per lane, logA from L2
        const size_t smem_g = (size_t)CPG_WARPS * 2 * r.p.Kp * sizeof(double);
        const int grid_g = (int)std::min<int64_t>((nseg + CPG_WARPS - 1) / CPG_WARPS, (int64_t)r.h->num_sms * 8);
        cp_sweep_generic_kernel<<<grid_g, 32 * CPG_WARPS, smem_g, r.st>>>(r.p, a);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        if (init_mode) CUDA_TRY(cudaMemsetAsync(r.d_counter, 0, sizeof(unsigned int), r.st));
        return CV_OK;
    }
    // one warp per segment (latency-oriented); the lock-step tile kernel only pays off with very many segments
    const size_t smem_c = (size_t)r.p.K * r.p.Kp * 8 + (r.p.bt_in_smem ? (size_t)r.p.M * r.p.Kp * 8 : 0) +
                          (size_t)CPW_WARPS * 2 * 16 * r.p.Kp;              // up to two segments per warp
    const bool fullwarp = g_tune.cp_fullwarp != 0;    // A/B hook: one segment per warp for every K
    const int64_t segs_per_block = (int64_t)CPW_WARPS * (r.p.Kp <= 16 && !fullwarp ? 2 : 1);
    const int grid = (int)std::min<int64_t>((ntask + segs_per_block - 1) / segs_per_block, (int64_t)r.h->num_sms * r.sweep_blocks_per_sm);
    const bool regs = r.p.Kp == 8 * ((r.p.K + 7) / 8);       // register-resident logA column needs Kp = 4 * KQ
    const int kq = regs ? (r.p.K + 7) / 8 : 9;
    if (fullwarp && kq == 1) cp_sweep_chain_kernel<1, 2><<<grid, 32 * CPW_WARPS, smem_c, r.st>>>(r.p, a);
    else if (fullwarp && kq == 2) cp_sweep_chain_kernel<1, 4><<<grid, 32 * CPW_WARPS, smem_c, r.st>>>(r.p, a);
    else switch (kq) {
        case 1: cp_sweep_chain_kernel<1, 2, 16><<<grid, 32 * CPW_WARPS, smem_c, r.st>>>(r.p, a); break;   // K <= 8
        case 2: cp_sweep_chain_kernel<1, 4, 16><<<grid, 32 * CPW_WARPS, smem_c, r.st>>>(r.p, a); break;   // K <= 16
        case 3: cp_sweep_chain_kernel<1, 6><<<grid, 32 * CPW_WARPS, smem_c, r.st>>>(r.p, a); break;
        case 4: cp_sweep_chain_kernel<1, 8><<<grid, 32 * CPW_WARPS, smem_c, r.st>>>(r.p, a); break;
        default:
            if (r.p.K <= 32) cp_sweep_chain_kernel<1, 0><<<grid, 32 * CPW_WARPS, smem_c, r.st>>>(r.p, a);
            else cp_sweep_chain_kernel<2, 0><<<grid, 32 * CPW_WARPS, smem_c, r.st>>>(r.p, a);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    if (init_mode) CUDA_TRY(cudaMemsetAsync(r.d_counter, 0, sizeof(unsigned int), r.st));   // nodes reset it in their sum kernel
    return CV_OK;
}

// cp.rs:85-93
int cp_backtrack(CpRun &r, double obj)
{
    if (!(obj > r.best_obj)) return fail(CV_ERR_ASSERT, "assert!(obj > self.best_obj) would fire (cp.rs:87)");
    r.best_obj = obj;
    const bool last = r.rank == r.R - 1;
    const int64_t rtop = last ? r.p.N - 1 : r.hi;                       // rows rtop .. lo+1 are walked by this rank
    const int nchunks = (int)std::max<int64_t>((rtop - r.lo + CP_BT_CHUNK - 1) / CP_BT_CHUNK, 1);
    double *d_obj = r.d_ub + 2;
    if (last) { cp_last_row_kernel<<<1, 32, 0, r.st>>>(r.p, d_obj, r.d_end); g_launches++; }
    const int K = r.p.K;
    const size_t rows_b = btr_rows_smem_bytes(K), chain_b = btr_chain_smem_bytes(nchunks, K);
    const int rows_in = rows_b <= BTR_SMEM_MAX ? 1 : 0, chain_in = chain_b <= BTR_SMEM_MAX ? 1 : 0;
    cp_btr_maps_kernel<<<nchunks, 64, rows_in ? rows_b : 0, r.st>>>(r.p, r.lo, rtop, nchunks, r.d_F, rows_in);
    const psi_t *maps = nullptr;
    int mapw = 0;
    PeerWait pw{nullptr, 1, 0u, nullptr};
    if (r.R > 1) {
        cv_cp_dist *dx = r.dx;
        mapw = dx->mapw;
        const size_t maps_off = dx->maps_off + (size_t)(r.n_map_x & 1) * CP_MAX_RANKS * mapw * sizeof(psi_t);
        r.n_map_x++;
        const unsigned int epoch = ++dx->epoch;
        cp_btr_total_kernel<<<1, 1024, chain_in ? chain_b : 0, r.st>>>(r.p, nchunks, r.d_F, r.d_end, dx->tab, maps_off, mapw, epoch, chain_in);
        g_launches++;
        pw = PeerWait{(const unsigned int *)dx->tab.base[r.rank], r.R, epoch, r.d_err};
        maps = (const psi_t *)(dx->tab.base[r.rank] + maps_off);
    }
    cp_btr_chain_kernel<<<1, 1024, chain_in ? chain_b : 0, r.st>>>(r.p, nchunks, r.d_F, r.d_end, maps, mapw, r.rank, r.R, r.d_entry,
                                                                 chain_in, pw);
    cp_btr_fill_kernel<<<nchunks, 64, rows_in ? rows_b : 0, r.st>>>(r.p, r.lo, rtop, r.hi, nchunks, r.d_entry, r.d_sol, rows_in);
    g_launches += 3;
    CUDA_TRY(cudaGetLastError());
    return CV_OK;
}

// ---- leaf batch ---------------------------------------------------------------------------------------------------
// The K nodes of the last component (cp.rs:96-124 with comp = ncomp-1) are leaves: each runs its sweeps, its bound
// and, if the bound beats best_obj, the backtrack; they all sweep the same segments and overwrite the same rows.
// Instead of K round trips of five dependent launches, their bounds are evaluated together from one batched sweep
// (last rows only), one batched terms launch and one batched exact-order sum; the host then walks the K bounds in
// the reference's order (explored_nodes, pruning test, best_obj) and replays at most two siblings on the real
// delta / psi state: the last one whose bound improved best_obj (its backtrack reads that state) and the last
// sibling evaluated (the state the reference leaves behind).  C2 backpointers of the other siblings -- one psi
// column each -- are written from the batch in sibling order around the replays, so every psi entry holds what
// the sequential loop would have left at that moment.  Bit-identical to the node-by-node loop (tests/test_cp_gpu.py
// runs both and compares the complete state).
int cp_leaf_group(CpRun &r, int32_t comp)
{
    const int K = r.p.K;
    const int64_t nseg = r.seg_off[comp + 1] - r.seg_off[comp];
    const int nterms = (int)r.cons_off[comp + 1];
    int nb = K;
    if (r.max_nodes) nb = (int)std::min<uint64_t>((uint64_t)K, r.max_nodes - r.explored);
    if (nb <= 0) return CV_OK;
    int rc;
    if ((rc = cp_sweep(r, r.seg_off[comp], nseg, 0, 0, nb))) return rc;
    CpLeafArgs a;
    a.term_pos = r.d_cons_pos; a.term_comp = r.d_term_comp; a.prev_seg = r.d_prev_seg; a.leaf_idx = r.d_leaf_idx;
    a.leaf_last = r.d_leaf_last; a.c2_out = r.d_c2; a.terms = r.d_leaf_terms; a.bsum = r.leaf_ws.bsum; a.bflag = r.leaf_ws.bflag;
    a.nterms = nterms; a.nseg = (int)nseg; a.nleaf = r.nleaf; a.last = comp;
    a.term_stride = r.leaf_term_stride; a.blk_stride = r.leaf_ws.blk_stride;
    cp_leaf_terms_kernel<<<dim3((unsigned)((nterms + SUM_BLK - 1) / SUM_BLK), (unsigned)nb), SUM_BLK, 0, r.st>>>(r.p, a);
    g_launches++;
    const UbSink sink{r.d_leaf_ub, nullptr, nullptr, 0ULL};
    if ((rc = cp_launch_sum(r.d_leaf_terms, nterms, sink, r.d_counter, r.leaf_ws, true, g_sum_force, r.st,
                            PeerWait{nullptr, 1, 0u, nullptr}, nb, r.leaf_term_stride)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(r.h_leaf_ub, r.d_leaf_ub, sizeof(double) * (size_t)nb, cudaMemcpyDeviceToHost, r.st));
    CUDA_TRY(cudaStreamSynchronize(r.st));
    // the reference's loop over the siblings, on the K bounds (cp.rs:96-123)
    int win = -1; double best = r.best_obj, win_ub = 0.0;
    for (int s = 0; s < nb; s++) {
        r.explored++;
        r.steps += r.seg_steps[comp];
        const double ub = r.h_leaf_ub[s];
        if (r.h->cp_ub.size() < (1u << 20)) r.h->cp_ub.push_back(ub);
        if (std::isnan(ub)) return fail(CV_ERR_NAN, "NaN upper bound");
        if (ub > best) { best = ub; win = s; win_ub = ub; }
    }
    const unsigned gl = (unsigned)((r.nleaf + 127) / 128);
    auto apply_c2 = [&](int s0, int s1) {
        if (s1 <= s0) return;
        cp_leaf_apply_c2_kernel<<<gl, 128, 0, r.st>>>(r.p, r.d_leaf_pos, r.nleaf, r.d_c2, s0, s1);
        g_launches++;
    };
    auto replay = [&](int s) -> int {                                  // the sibling's viterbi_from calls on the real state
        int rc2 = cp_sweep(r, r.seg_off[comp], nseg, s, 0);
        if (rc2) return rc2;
        const int64_t nfix = r.fix_off[comp + 1] - r.fix_off[comp];
        cp_fixup_kernel<<<(unsigned)((std::max<int64_t>(nfix, 1) + 127) / 128), 128, 0, r.st>>>(
            r.p, r.d_fix_pos + r.fix_off[comp], (int)nfix, comp, s, r.lo, r.hi);
        g_launches++;
        return CV_OK;
    };
    int done = 0;                                                      // siblings whose psi columns are in place
    if (win >= 0) {
        apply_c2(0, win);
        if ((rc = replay(win))) return rc;
        if ((rc = cp_backtrack(r, win_ub))) return rc;                 // cp.rs:121
        done = win + 1;
    }
    if (win != nb - 1) {
        apply_c2(done, nb - 1);
        if ((rc = replay(nb - 1))) return rc;
    }
    CUDA_TRY(cudaGetLastError());
    return CV_OK;
}

// cp.rs:95-126
int cp_solve_r(CpRun &r, int32_t comp)
{
    const int K = r.p.K;
    const int64_t nseg = r.seg_off[comp + 1] - r.seg_off[comp];        // this rank's sweeps of the component
    const int64_t nfix = r.fix_off[comp + 1] - r.fix_off[comp];
    const int nterms = (int)r.cons_off[comp + 1];
    const int nlocal = (int)r.lterm_off[comp + 1];
    if (r.leaf_batch && comp + 1 == r.ncomp && nseg > 0) return cp_leaf_group(r, comp);
    for (int state = 0; state < K; state++) {
        if (r.max_nodes && r.explored >= r.max_nodes) break;            // builder-added, deterministic budget
        r.explored++;                                                    // cp.rs:97
        const bool prof = g_tune.cp_prof != 0;
        auto tick = [&](int slot, std::chrono::steady_clock::time_point &t0) {
            if (!prof) return;
            cudaStreamSynchronize(r.st);
            auto t1 = std::chrono::steady_clock::now();
            g_cp_prof[slot] += std::chrono::duration<double, std::micro>(t1 - t0).count();
            t0 = t1;
        };
        auto t0 = std::chrono::steady_clock::now();
        int rc = cp_sweep(r, r.seg_off[comp], nseg, state, 0);          // cp.rs:99-102, phases A + B
        if (rc) return rc;
        tick(0, t0);
        r.steps += r.seg_steps[comp];
        cp_fixup_kernel<<<(unsigned)((std::max<int64_t>(nfix, 1) + 127) / 128), 128, 0, r.st>>>(
            r.p, r.d_fix_pos + r.fix_off[comp], (int)nfix, comp, state, r.lo, r.hi);   // also records cstr_choices[comp] (cp.rs:98)
        g_launches++;
        tick(1, t0);
        bool have_stats = nterms > 0;
        PeerWait pw{nullptr, 1, 0u, nullptr};
        if (r.R > 1) {
            // the rank's terms go straight into every rank's term list (peer stores + flag), then wait for all ranks
            cv_cp_dist *dx = r.dx;
            const size_t toff = dx->terms_off + (size_t)(r.n_term_x & 1) * dx->cap_terms * sizeof(double);
            r.n_term_x++;
            const unsigned int epoch = ++dx->epoch;
            cp_terms_peer_kernel<<<std::max((nlocal + SUM_BLK - 1) / SUM_BLK, 1), SUM_BLK, 0, r.st>>>(
                r.p, r.d_lterm_pos, r.d_lterm_comp, r.d_lterm_gidx, nlocal, dx->tab, toff, epoch, r.d_done);
            g_launches++;
            r.d_terms = (double *)(dx->tab.base[r.rank] + toff);
            have_stats = false;
            pw = PeerWait{(const unsigned int *)dx->tab.base[r.rank], r.R, epoch, r.d_err};
        } else if (nterms > 0) {
            cp_terms_kernel<<<(nterms + SUM_BLK - 1) / SUM_BLK, SUM_BLK, 0, r.st>>>(r.p, r.d_cons_pos, r.d_term_comp, nterms,
                                                                                     r.d_terms, r.sum_ws.bsum, r.sum_ws.bflag);
            g_launches++;
        }
        tick(2, t0);
        // cp.rs:103-116: the exact-order sum (parallel binade scan for long lists, plain loop for short ones)
        const unsigned long long seq = ++r.node_seq;
        const UbSink sink{r.d_ub, r.poll ? r.h_ub : nullptr, r.R > 1 ? r.d_err : nullptr, seq};
        if ((rc = cp_launch_sum(r.d_terms, nterms, sink, r.d_counter, r.sum_ws, have_stats, g_sum_force, r.st, pw))) return rc;
        tick(3, t0);
        if (r.poll) {
            // the sum kernel writes ub + the node's sequence number into pinned host memory; poll for it
            volatile unsigned long long *seqp = reinterpret_cast<volatile unsigned long long *>(r.h_ub + 2);
            for (unsigned int spin = 1; *seqp != seq; spin++) {
                if ((spin & 0x3fff) == 0) {
                    const cudaError_t e = cudaStreamQuery(r.st);
                    if (e == cudaSuccess) { if (*seqp == seq) break; return fail(CV_ERR_CUDA, "bound of node %llu never arrived", seq); }
                    if (e != cudaErrorNotReady) return fail(CV_ERR_CUDA, "stream error while waiting for a node's bound: %s", cudaGetErrorString(e));
                }
            }
        } else {
            // (the word after ub is the peer-timeout flag of the wait kernels, always 0.0 on a single rank)
            CUDA_TRY(cudaMemcpyAsync(r.h_ub, r.d_ub, (r.R > 1 ? 2 : 1) * sizeof(double), cudaMemcpyDeviceToHost, r.st));
            CUDA_TRY(cudaStreamSynchronize(r.st));
        }
        tick(4, t0);
        const double ub = *r.h_ub;
        if (r.R > 1 && r.h_ub[1] != 0.0) return fail(CV_ERR_CUDA, "a peer rank did not publish its bound terms within the time limit");
        if (r.h->cp_ub.size() < (1u << 20)) r.h->cp_ub.push_back(ub);
        if (std::isnan(ub)) return fail(CV_ERR_NAN, "NaN upper bound");
        if (ub > r.best_obj) {                                           // cp.rs:117
            if (comp + 1 < r.ncomp) {                                    // cp.rs:118-119
                if ((rc = cp_solve_r(r, comp + 1))) return rc;
            } else {
                if ((rc = cp_backtrack(r, ub))) return rc;               // cp.rs:121
            }
        }
    }
    // cp.rs:125 (cstr_choices[comp] = None): which positions count as fixed is a function of the depth alone
    // (segments are precomputed per depth) and choice[comp] is rewritten before it is read again.
    return CV_OK;
}

template <class T>
int upload(DevBuf &b, const std::vector<T> &v, T **out, cudaStream_t st)
{
    int rc = b.ensure(std::max<size_t>(sizeof(T) * v.size(), 16));
    if (rc) return rc;
    if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(b.p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
    *out = (T *)b.p;
    return CV_OK;
}

}  // namespace

extern "C" int cv_cp_dist_create(cv_hmm *h, int rank, int nranks, int64_t cap_N, int64_t cap_terms, void *handle_out,
                                 cv_cp_dist **out)
{
    if (!h || !out || !handle_out) return fail(CV_ERR_ARG, "NULL argument");
    if (nranks < 1 || nranks > CP_MAX_RANKS || rank < 0 || rank >= nranks)
        return fail(CV_ERR_ARG, "rank %d / nranks %d outside 1..%d", rank, nranks, CP_MAX_RANKS);
    if (cap_N <= 0 || cap_terms < 0) return fail(CV_ERR_ARG, "bad capacity");
    static_assert(sizeof(cudaIpcMemHandle_t) == CV_IPC_HANDLE_BYTES, "cv_b200.h: CV_IPC_HANDLE_BYTES");
    CUDA_TRY(cudaSetDevice(h->device));
    cv_cp_dist *d = new cv_cp_dist;
    d->h = h; d->rank = rank; d->R = nranks; d->cap_N = cap_N; d->cap_terms = cap_terms;
    d->mapw = (h->K + 1 + 7) / 8 * 8;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    d->maps_off = CP_FLAGS_BYTES;
    d->terms_off = up(d->maps_off + (size_t)2 * CP_MAX_RANKS * d->mapw * sizeof(psi_t));
    d->sol_off = up(d->terms_off + (size_t)2 * (cap_terms + 8) * sizeof(double));
    d->bytes = up(d->sol_off + (size_t)cap_N * sizeof(uint64_t));
    cudaError_t e = cudaMalloc(&d->mine, d->bytes);
    if (e != cudaSuccess) { delete d; cudaGetLastError(); return fail(CV_ERR_OOM, "cudaMalloc(%zu) failed: %s", d->bytes, cudaGetErrorString(e)); }
    e = cudaMemset(d->mine, 0, d->bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out, d->mine);
    if (e != cudaSuccess) { cudaFree(d->mine); delete d; cudaGetLastError(); return fail(CV_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    d->tab.R = nranks; d->tab.rank = rank;
    d->tab.base[rank] = (unsigned char *)d->mine;
    *out = d;
    return CV_OK;
}

extern "C" int cv_cp_dist_connect(cv_cp_dist *d, const void *all_handles)
{
    if (!d || !all_handles) return fail(CV_ERR_ARG, "NULL argument");
    if (d->connected) return CV_OK;
    CUDA_TRY(cudaSetDevice(d->h->device));
    const cudaIpcMemHandle_t *hs = (const cudaIpcMemHandle_t *)all_handles;
    for (int q = 0; q < d->R; q++) {
        if (q == d->rank) continue;
        void *ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, hs[q], cudaIpcMemLazyEnablePeerAccess));
        d->tab.base[q] = (unsigned char *)ptr;
    }
    d->connected = true;
    return CV_OK;
}

extern "C" void cv_cp_dist_destroy(cv_cp_dist *d)
{
    if (!d) return;
    cudaSetDevice(d->h->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < d->R; q++)
        if (q != d->rank && d->tab.base[q]) cudaIpcCloseMemHandle(d->tab.base[q]);
    if (d->mine) cudaFree(d->mine);
    delete d;
}

// The row cuts of the sharded solve: cuts[0] = 0 < cuts[1] < .. < cuts[R] = N, inner cuts are positions of
// component 0 closest to an even split.  Returns false when component 0 has too few positions (then every
// rank solves the whole problem itself: "replicas").  Pure host logic, also exported for the CPU tests.
static bool cp_plan_cuts(const int32_t *comp, int64_t N, int R, std::vector<int64_t> &cuts)
{
    cuts.assign((size_t)R + 1, 0);
    cuts[R] = N;
    if (R == 1) return true;
    std::vector<int64_t> c0;
    for (int64_t t = 1; t < N; t++) if (comp[t] == 0) c0.push_back(t);
    for (int r = 1; r < R; r++) {
        const int64_t want = N / R * r + (N % R) * r / R;
        auto it = std::lower_bound(c0.begin(), c0.end(), want);
        int64_t best = -1;
        if (it != c0.end()) best = *it;
        if (it != c0.begin() && (best < 0 || want - *(it - 1) <= best - want)) best = *(it - 1);
        if (best <= cuts[r - 1]) {                                     // keep the cuts strictly increasing
            auto nx = std::upper_bound(c0.begin(), c0.end(), cuts[r - 1]);
            if (nx == c0.end()) return false;
            best = *nx;
        }
        cuts[r] = best;
    }
    for (int r = 1; r <= R; r++) if (cuts[r] <= cuts[r - 1]) return false;
    return true;
}

extern "C" int cv_cp_plan_cuts(const int32_t *comp, int64_t N, int nranks, int64_t *cuts_out)
{
    if (!comp || !cuts_out || N <= 0 || nranks < 1) return fail(CV_ERR_ARG, "bad argument");
    std::vector<int64_t> cuts;
    const bool ok = cp_plan_cuts(comp, N, nranks, cuts);
    if (!ok) { for (int r = 0; r <= nranks; r++) cuts_out[r] = r == 0 ? 0 : N; return CV_OK; }   // replicas: rank 0 owns [0, N)
    for (int r = 0; r <= nranks; r++) cuts_out[r] = cuts[r];
    return CV_OK;
}

static int cp_solve_impl(cv_hmm *h, cv_cp_dist *dx, const uint32_t *obs, const uint8_t *is_seq_start, const int32_t *comp,
                         int64_t N, int32_t ncomp, uint64_t max_nodes, uint64_t *sol_out, double *obj_out,
                         uint64_t *explored_out, uint64_t *steps_out)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    if (N <= 0) return fail(CV_ERR_EMPTY, "empty super-sequence (reference: array.row(len-1) panics)");
    if (!obs || !is_seq_start || !comp || !sol_out) return fail(CV_ERR_ARG, "NULL buffer");
    if (ncomp < 0) return fail(CV_ERR_ARG, "ncomp < 0");
    if ((size_t)4 * 2 * (h->K > SMALL_K_MAX ? h->Kl : h->Kp) * sizeof(double) > 200 * 1024)
        return fail(CV_ERR_UNSUPPORTED, "cv_cp_solve: K = %d needs more shared memory than an SM has", h->K);
    for (int64_t t = 0; t < N; t++) {
        if ((int64_t)obs[t] >= h->M) return fail(CV_ERR_ARG, "observation index >= M at %lld (reference: ndarray index panic)", (long long)t);
        if (comp[t] >= ncomp) return fail(CV_ERR_ARG, "component id %d >= ncomp %d at %lld (reference: constraints[ucomp] index panic, cp.rs:25)", comp[t], ncomp, (long long)t);
    }
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int K = h->K;
    h->cp_ub.clear();
    h->cp_N = 0;

    CpRun r;
    r.h = h; r.st = st; r.ncomp = ncomp; r.max_nodes = max_nodes;

    // ---- sharding: row cuts at positions of component 0 (cp_dist.cuh); too few of them => replicas ----
    std::vector<int64_t> cuts;
    r.lo = 0; r.hi = N;
    if (dx && dx->R > 1) {
        if (dx->h != h) return fail(CV_ERR_ARG, "exchange buffer belongs to another model handle");
        if (!dx->connected) return fail(CV_ERR_ARG, "cv_cp_dist_connect has not run");
        if (h->K + 1 > 1024) return fail(CV_ERR_UNSUPPORTED, "sharded constrained decode supports K <= 1023");
        if (N > dx->cap_N) return fail(CV_ERR_ARG, "N = %lld exceeds the exchange buffer's capacity %lld", (long long)N, (long long)dx->cap_N);
        if (ncomp > 0 && cp_plan_cuts(comp, N, dx->R, cuts)) {
            r.dx = dx; r.R = dx->R; r.rank = dx->rank;
            r.lo = cuts[r.rank]; r.hi = cuts[r.rank + 1];
        }
    }
    const bool sharded = r.R > 1;
    const int64_t lo = r.lo, hi = r.hi;

    // ---- CPSolver::new (cp.rs:20-30): positions of every component, ascending ----
    std::vector<std::vector<int64_t>> cons((size_t)ncomp);
    for (int64_t t = 0; t < N; t++) if (comp[t] >= 0) cons[comp[t]].push_back(t);
    std::vector<int64_t> cons_pos; std::vector<int32_t> term_comp;
    r.cons_off.assign((size_t)ncomp + 1, 0);
    for (int32_t c = 0; c < ncomp; c++) {
        r.cons_off[c] = (int64_t)cons_pos.size();
        for (int64_t t : cons[c]) { cons_pos.push_back(t); term_comp.push_back(c); }
    }
    r.cons_off[ncomp] = (int64_t)cons_pos.size();
    if (cons_pos.size() > 0x7fffffffULL) return fail(CV_ERR_UNSUPPORTED, "too many clamped positions");
    if (sharded && (int64_t)cons_pos.size() > dx->cap_terms)
        return fail(CV_ERR_ARG, "%zu clamped positions exceed the exchange buffer's capacity %lld", cons_pos.size(), (long long)dx->cap_terms);

    // fix-up positions of this rank per depth ([lo, hi]: C1 for pos < hi, C2 for pos > lo) and its bound terms
    // (t in (lo, hi], t = 0 on the first rank) with their index in the reference's summation order
    std::vector<int64_t> fix_pos, lterm_pos; std::vector<int32_t> lterm_comp, lterm_gidx;
    r.fix_off.assign((size_t)ncomp + 1, 0); r.lterm_off.assign((size_t)ncomp + 1, 0);
    for (int32_t c = 0; c < ncomp; c++) {
        r.fix_off[c] = (int64_t)fix_pos.size(); r.lterm_off[c] = (int64_t)lterm_pos.size();
        for (size_t i = 0; i < cons[c].size(); i++) {
            const int64_t t = cons[c][i];
            if (t >= lo && t <= hi) fix_pos.push_back(t);
            if ((t > lo && t <= hi) || (t == 0 && lo == 0)) {
                lterm_pos.push_back(t); lterm_comp.push_back(c); lterm_gidx.push_back((int32_t)(r.cons_off[c] + (int64_t)i));
            }
        }
    }
    r.fix_off[ncomp] = (int64_t)fix_pos.size(); r.lterm_off[ncomp] = (int64_t)lterm_pos.size();

    // ---- sweep segments ----
    // depth c (components 0..c assigned): one sweep per position of c, running to the next position whose
    // component is <= c (cp.rs:48); segment 0 of the list is the init_viterbi prefix (cp.rs:63-83).  A rank
    // keeps the sweeps that start in its rows; seg_steps counts every rank's rows (the reported work).
    std::vector<int64_t> seg_from; std::vector<int32_t> seg_len;
    int64_t prefix = 0;
    while (prefix < N && comp[prefix] < 0) prefix++;                 // rows 0 .. prefix-1 are swept by init_viterbi
    seg_from.push_back(0); seg_len.push_back((int32_t)std::max<int64_t>(prefix - 1, 0));
    r.seg_off.assign((size_t)ncomp + 1, 1);
    r.seg_steps.assign((size_t)std::max(ncomp, 1), 0);
    {
        std::vector<int64_t> next_fixed((size_t)N + 1);
        for (int32_t c = 0; c < ncomp; c++) {
            r.seg_off[c] = (int64_t)seg_from.size();
            int64_t nf = N;
            for (int64_t t = N - 1; t >= 0; t--) { next_fixed[t] = nf; if (comp[t] >= 0 && comp[t] <= c) nf = t; }
            std::vector<std::pair<int64_t, int64_t>> segs;           // (len, from)
            for (int64_t pos : cons[c]) {
                const int64_t len = next_fixed[pos] - pos - 1;
                if (len > 0x7fffffffLL) return fail(CV_ERR_UNSUPPORTED, "segment too long");
                r.seg_steps[c] += (uint64_t)len;
                if (pos >= lo && pos < hi) segs.emplace_back(len, pos);
            }
            std::stable_sort(segs.begin(), segs.end(), [](const auto &x, const auto &y) { return x.first > y.first; });
            for (auto &s : segs) { seg_from.push_back(s.second); seg_len.push_back((int32_t)s.first); }
        }
        r.seg_off[ncomp] = (int64_t)seg_from.size();
    }

    // ---- device state: delta / psi rows [lo, hi] of this rank (indexed by global row through a shifted base) ----
    DevBuf *b = h->cpb;
    int rc;
    const int64_t nrows = std::min<int64_t>(hi + 1, N) - lo;
    if ((rc = b[0].ensure(sizeof(double) * (size_t)nrows * K))) return rc;
    if ((rc = b[1].ensure(sizeof(psi_t) * (size_t)nrows * K))) return rc;
    CUDA_TRY(cudaMemsetAsync(b[0].p, 0, sizeof(double) * (size_t)nrows * K, st));     // Array2::from_elem(.., 0.0) cp.rs:134
    CUDA_TRY(cudaMemsetAsync(b[1].p, 0, sizeof(psi_t) * (size_t)nrows * K, st));      // bt = 0                    cp.rs:135
    if ((rc = b[2].ensure(sizeof(uint32_t) * (size_t)N))) return rc;
    if ((rc = b[3].ensure((size_t)N))) return rc;
    if ((rc = b[4].ensure(sizeof(int32_t) * (size_t)N))) return rc;
    CUDA_TRY(cudaMemcpyAsync(b[2].p, obs, sizeof(uint32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b[3].p, is_seq_start, (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b[4].p, comp, sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
    std::vector<int32_t> choice0((size_t)std::max(ncomp, 1), -1);
    int32_t *d_choice = nullptr;
    if ((rc = upload(b[5], choice0, &d_choice, st))) return rc;
    if ((rc = upload(b[6], seg_from, &r.d_seg_from, st))) return rc;
    if ((rc = upload(b[7], seg_len, &r.d_seg_len, st))) return rc;
    if (sharded) {
        // one upload: [fix_pos | lterm_pos] as i64, [lterm_comp | lterm_gidx] as i32
        std::vector<int64_t> v64(fix_pos); v64.insert(v64.end(), lterm_pos.begin(), lterm_pos.end());
        std::vector<int32_t> v32(lterm_comp); v32.insert(v32.end(), lterm_gidx.begin(), lterm_gidx.end());
        int64_t *d64 = nullptr; int32_t *d32 = nullptr;
        if ((rc = upload(b[8], v64, &d64, st))) return rc;
        if ((rc = upload(b[9], v32, &d32, st))) return rc;
        r.d_fix_pos = d64; r.d_lterm_pos = d64 + fix_pos.size();
        r.d_lterm_comp = d32; r.d_lterm_gidx = d32 + lterm_comp.size();
    } else {
        if ((rc = upload(b[8], cons_pos, &r.d_cons_pos, st))) return rc;
        if ((rc = upload(b[9], term_comp, &r.d_term_comp, st))) return rc;
        r.d_fix_pos = r.d_cons_pos;                                                // [lo, hi] = everything
    }
    if ((rc = b[10].ensure(sizeof(double) * (cons_pos.size() + 16) + 256))) return rc;
    // scalars: [0] ub  [1] peer-timeout flag  [2] obj  [3] end state  [4] segment counter | publish counter
    double *sc = (double *)b[10].p;
    r.d_ub = sc; r.d_err = (int *)(sc + 1); r.d_end = (int *)(sc + 3);
    r.d_counter = (unsigned int *)(sc + 4); r.d_done = r.d_counter + 1;
    r.d_terms = sc + 8;
    CUDA_TRY(cudaMemsetAsync(b[10].p, 0, 64, st));
    const int64_t walk_rows = (r.rank == r.R - 1 ? N - 1 : hi) - lo;
    const int nchunks = (int)std::max<int64_t>((walk_rows + CP_BT_CHUNK - 1) / CP_BT_CHUNK, 1);
    if ((rc = b[11].ensure(sizeof(psi_t) * (size_t)nchunks * K + 64))) return rc;
    if ((rc = b[12].ensure(sizeof(int) * (size_t)nchunks + 64))) return rc;
    r.d_F = (psi_t *)b[11].p; r.d_entry = (int *)b[12].p;
    if (sharded) {
        r.d_sol = (uint64_t *)(dx->tab.base[r.rank] + dx->sol_off);                // own rows here, peers' rows arrive at the end
    } else {
        if ((rc = b[13].ensure(sizeof(uint64_t) * (size_t)N))) return rc;
        r.d_sol = (uint64_t *)b[13].p;
    }
    CUDA_TRY(cudaMemsetAsync(r.d_sol, 0, sizeof(uint64_t) * (size_t)N, st));       // best_sol = 0 (cp.rs:29)
    if ((rc = r.sum_ws.bind(b[14], cons_pos.size()))) return rc;
    // ---- leaf batch: scratch for the K siblings of the last component (see cp_leaf_group) ----
    {
        const int32_t last = ncomp - 1;
        const int64_t nleaf = ncomp > 0 ? r.seg_off[ncomp] - r.seg_off[last] : 0;
        const size_t scratch = (size_t)K * (size_t)nleaf * K * sizeof(double) + (size_t)K * (cons_pos.size() + 1) * sizeof(double);
        r.leaf_batch = g_tune.cp_leaf_batch && !sharded && K <= SMALL_K_MAX && K >= 2 && nleaf > 0 &&
                       nleaf * (int64_t)K < 0x7fffffffLL && scratch < ((size_t)4 << 30);
        if (r.leaf_batch) {
            const int64_t *lf = seg_from.data() + r.seg_off[last];             // the last component's sweeps, device order
            std::vector<int32_t> seg_of_pos((size_t)N, -1);
            for (int64_t g = 0; g < nleaf; g++) seg_of_pos[(size_t)lf[g]] = (int32_t)g;
            // nearest clamped position at or below every row
            std::vector<int32_t> prev_seg(cons_pos.size(), -1), leaf_idx(cons_pos.size(), -1);
            std::vector<int64_t> below((size_t)N, -1);
            int64_t lastc = -1;
            for (int64_t t = 0; t < N; t++) { if (comp[t] >= 0) lastc = t; below[(size_t)t] = lastc; }
            for (size_t k = 0; k < cons_pos.size(); k++) {
                const int64_t t = cons_pos[k];
                if (term_comp[k] == last) leaf_idx[k] = seg_of_pos[(size_t)t];
                if (t > 0) {
                    const int64_t q = below[(size_t)(t - 1)];
                    if (q >= 0 && comp[q] == last) prev_seg[k] = seg_of_pos[(size_t)q];
                }
            }
            std::vector<int32_t> both(prev_seg); both.insert(both.end(), leaf_idx.begin(), leaf_idx.end());
            int32_t *d32 = nullptr;
            if ((rc = upload(b[15], both, &d32, st))) return rc;
            r.d_prev_seg = d32; r.d_leaf_idx = d32 + prev_seg.size();
            r.nleaf = (int)nleaf;
            r.d_leaf_pos = r.d_seg_from + r.seg_off[last];
            r.leaf_term_stride = (long long)((cons_pos.size() + SUM_BLK) / SUM_BLK * SUM_BLK);
            if ((rc = b[16].ensure(sizeof(double) * (size_t)K * (size_t)nleaf * K))) return rc;
            if ((rc = b[17].ensure(sizeof(psi_t) * (size_t)K * (size_t)nleaf + 64))) return rc;
            if ((rc = b[18].ensure(sizeof(double) * ((size_t)K * (size_t)r.leaf_term_stride + K) + 64))) return rc;
            r.d_leaf_last = (double *)b[16].p; r.d_c2 = (psi_t *)b[17].p;
            r.d_leaf_terms = (double *)b[18].p; r.d_leaf_ub = r.d_leaf_terms + (size_t)K * (size_t)r.leaf_term_stride;
            if ((rc = r.leaf_ws.bind(b[19], cons_pos.size(), (size_t)K))) return rc;
            r.h_leaf_ub = (double *)((char *)h->pinned_status + 1024);
        }
    }
    g_sum_force = g_tune.cp_sum;
    r.h_ub = (double *)h->pinned_status + 1;
    r.h_ub[0] = r.h_ub[1] = r.h_ub[2] = 0.0;
    r.poll = g_tune.cp_hostpoll != 0;

    CpParams &p = r.p;
    const bool small = K <= SMALL_K_MAX;
    p.A = small ? h->dA : h->dAl; p.BT = small ? h->dBT : h->dBTl; p.Pi = h->dPi;
    p.obs = (const uint32_t *)b[2].p; p.start = (const uint8_t *)b[3].p; p.comp = (const int32_t *)b[4].p;
    p.delta = (double *)b[0].p - (size_t)lo * K; p.psi = (psi_t *)b[1].p - (size_t)lo * K; p.choice = d_choice;
    p.N = N; p.M = h->M; p.K = K; p.Kp = small ? h->Kp : h->Kl; p.G = h->G;
    p.bt_in_smem = (small && (size_t)h->M * h->Kp * 8 <= CHAIN_BT_SMEM_MAX) ? 1 : 0;
    if (!small) {
        const size_t smem_g = (size_t)CPG_WARPS * 2 * p.Kp * sizeof(double);
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    }
    if (small) {
        const size_t smem_c = (size_t)K * h->Kp * 8 + (p.bt_in_smem ? (size_t)h->M * h->Kp * 8 : 0) + (size_t)CPW_WARPS * 2 * 16 * h->Kp;
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_chain_kernel<1, 2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_chain_kernel<1, 4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_chain_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_chain_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_chain_kernel<1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_chain_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_chain_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        CUDA_TRY(cudaFuncSetAttribute(cp_sweep_chain_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        // segments are dealt statically to the resident warps: one wave of CTAs
        int nb = 0;
        const bool regs = h->Kp == 8 * ((K + 7) / 8);
        const int kq = regs ? (K + 7) / 8 : 9;
        cudaError_t oe;
        if (kq == 1) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cp_sweep_chain_kernel<1, 2, 16>, 32 * CPW_WARPS, smem_c);
        else if (kq == 2) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cp_sweep_chain_kernel<1, 4, 16>, 32 * CPW_WARPS, smem_c);
        else if (kq == 3) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cp_sweep_chain_kernel<1, 6>, 32 * CPW_WARPS, smem_c);
        else if (kq == 4) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cp_sweep_chain_kernel<1, 8>, 32 * CPW_WARPS, smem_c);
        else if (K <= 32) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cp_sweep_chain_kernel<1, 0>, 32 * CPW_WARPS, smem_c);
        else oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cp_sweep_chain_kernel<2, 0>, 32 * CPW_WARPS, smem_c);
        if (oe != cudaSuccess) { cudaGetLastError(); nb = 0; }
        r.sweep_blocks_per_sm = std::max(1, std::min(nb, 16));
    }

    CUDA_TRY(cudaFuncSetAttribute(cp_sum_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sum_chain_smem_bytes(SUM_MAX_BLOCKS)));
    CUDA_TRY(cudaFuncSetAttribute(cp_btr_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BTR_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(cp_btr_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BTR_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(cp_btr_total_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BTR_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(cp_btr_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BTR_SMEM_MAX));

    const bool timing = g_timing.load() != 0;
    if (timing) CUDA_TRY(cudaEventRecord(h->ev0, st));
    // ---- init_viterbi (cp.rs:63-83): the prefix lies before the first cut, so it is the first rank's ----
    if (prefix >= 1 && lo == 0) {
        cp_row0_kernel<<<(K + 255) / 256, 256, 0, st>>>(p);
        g_launches++;
        if (prefix > 1) {
            if ((rc = cp_sweep(r, 0, 1, 0, 1))) return rc;
        }
    }
    if (prefix > 1) r.steps += (uint64_t)(prefix - 1);
    // ---- solve (cp.rs:137-142) ----
    if (ncomp > 0) {
        rc = cp_solve_r(r, 0);
    } else {
        double *d_obj = r.d_ub + 2;
        cp_last_row_kernel<<<1, 32, 0, st>>>(p, d_obj, r.d_end);
        g_launches++;
        CUDA_TRY(cudaMemcpyAsync(r.h_ub, d_obj, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        rc = cp_backtrack(r, *r.h_ub);
    }
    if (rc) return rc;
    if (sharded) {
        // every rank's rows of best_sol into every rank's buffer (peer stores + flag), then wait for all of them
        const unsigned int epoch = ++dx->epoch;
        const int grid = (int)std::min<int64_t>((hi - lo + 255) / 256, (int64_t)h->num_sms * 4);
        cp_sol_publish_kernel<<<std::max(grid, 1), 256, 0, st>>>(r.d_sol, lo, hi, dx->tab, dx->sol_off, epoch, r.d_done);
        cp_peer_wait_kernel<<<1, 32, 0, st>>>(PeerWait{(const unsigned int *)dx->tab.base[r.rank], r.R, epoch, r.d_err});
        g_launches += 2;
        CUDA_TRY(cudaMemcpyAsync(r.h_ub + 1, r.d_err, sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (timing) {
        CUDA_TRY(cudaEventRecord(h->ev1, st));
        CUDA_TRY(cudaEventRecord(h->ev2, st));
    }
    h->cp_N = sharded ? 0 : N;                                       // the state hooks read a whole-problem delta / psi
    CUDA_TRY(cudaMemcpyAsync(sol_out, r.d_sol, sizeof(uint64_t) * (size_t)N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (sharded && r.h_ub[1] != 0.0) return fail(CV_ERR_CUDA, "a peer rank did not publish its part of the solution within the time limit");
    if (timing) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_ms = ms; else cudaGetLastError();
        h->last_bt_ms = 0.0;
    }
    if (g_tune.cp_prof) {
        fprintf(stderr, "[cv] cp phases (us/node over %llu nodes): sweep %.1f fixup %.1f terms %.1f sum %.1f readback %.1f\n",
                (unsigned long long)r.explored, g_cp_prof[0] / r.explored, g_cp_prof[1] / r.explored,
                g_cp_prof[2] / r.explored, g_cp_prof[3] / r.explored, g_cp_prof[4] / r.explored);
        for (double &v : g_cp_prof) v = 0.0;
    }
    if (obj_out) *obj_out = r.best_obj;
    if (explored_out) *explored_out = r.explored;
    if (steps_out) *steps_out = r.steps;
    return rc;
}

extern "C" int cv_cp_solve(cv_hmm *h, const uint32_t *obs, const uint8_t *is_seq_start, const int32_t *comp, int64_t N,
                           int32_t ncomp, uint64_t max_nodes, uint64_t *sol_out, double *obj_out,
                           uint64_t *explored_out, uint64_t *steps_out)
{
    return cp_solve_impl(h, nullptr, obs, is_seq_start, comp, N, ncomp, max_nodes, sol_out, obj_out, explored_out, steps_out);
}

extern "C" int cv_cp_solve_dist(cv_cp_dist *d, const uint32_t *obs, const uint8_t *is_seq_start, const int32_t *comp, int64_t N,
                                int32_t ncomp, uint64_t max_nodes, uint64_t *sol_out, double *obj_out,
                                uint64_t *explored_out, uint64_t *steps_out)
{
    if (!d) return fail(CV_ERR_ARG, "NULL exchange buffer");
    return cp_solve_impl(d->h, d, obs, is_seq_start, comp, N, ncomp, max_nodes, sol_out, obj_out, explored_out, steps_out);
}

extern "C" int cv_debug_cp_last_state(cv_hmm *h, double *delta_out, uint64_t *psi_out)
{
    if (!h || h->cp_N <= 0) return fail(CV_ERR_ARG, "no constrained solve has run on this model");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t n = (size_t)h->cp_N * h->K;
    if (delta_out) CUDA_TRY(cudaMemcpy(delta_out, h->cpb[0].p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (psi_out) {
        uint64_t *tmp = nullptr;
        CUDA_TRY(cudaMalloc(&tmp, sizeof(uint64_t) * n));
        cp_widen_psi_kernel<<<(unsigned)((n + 255) / 256), 256>>>((const psi_t *)h->cpb[1].p, (int64_t)n, tmp);
        g_launches++;
        CUDA_TRY(cudaMemcpy(psi_out, tmp, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
        cudaFree(tmp);
    }
    return CV_OK;
}

extern "C" int cv_debug_cp_last_ub(cv_hmm *h, double *ub_out, uint64_t cap, uint64_t *n_out)
{
    if (!h) return fail(CV_ERR_ARG, "NULL model");
    const uint64_t n = std::min<uint64_t>(cap, h->cp_ub.size());
    if (ub_out) for (uint64_t i = 0; i < n; i++) ub_out[i] = h->cp_ub[i];
    if (n_out) *n_out = h->cp_ub.size();
    return CV_OK;
}

// Debug/parity hook: the ordered sum of `n` host values with the serial (mode 0) or the parallel (mode 1) kernel.
extern "C" int cv_debug_ordered_sum(const double *values, int64_t n, int mode, double *out)
{
    int rc = check_device(0);
    if (rc) return rc;
    if (n < 0 || n > 0x7fffffffLL || !out) return fail(CV_ERR_ARG, "bad argument");
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(double) * (size_t)(n + 1)));
    if (n) CUDA_TRY(cudaMemcpy(d + 1, values, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    DevBuf wsb; SumWs ws;
    if ((rc = ws.bind(wsb, (size_t)n))) { cudaFree(d); return rc; }
    if (mode < 0 || mode > 2) mode = 0;
    if (cudaFuncSetAttribute(cp_sum_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sum_chain_smem_bytes(SUM_MAX_BLOCKS)) != cudaSuccess) { cudaFree(d); wsb.release(); return fail(CV_ERR_CUDA, "cudaFuncSetAttribute"); }
    rc = cp_launch_sum(d + 1, (int)n, UbSink{d, nullptr, nullptr, 0ULL}, nullptr, ws, false, mode, nullptr, PeerWait{nullptr, 1, 0u, nullptr});
    if (rc) { cudaFree(d); wsb.release(); return rc; }
    CUDA_TRY(cudaMemcpy(out, d, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d);
    wsb.release();
    return CV_OK;
}

// net_kernel.cuh — hand-written template of the *network* execution model.
//
// One persistent cooperative grid (one CTA per SM, PRB_BLOCK threads) integrates a whole network whose
// state is too large for one thread: populations of N units with dense / pair-space / gather coupling.
// The time loop, all Runge-Kutta stages and the observer store run inside ONE launch; CTAs meet at a
// grid-wide barrier after every rhs evaluation (the coupling needs the complete stage vector) and at the
// few extra points where the generated code consumes a vector at another index than it was produced.
//
// Replaces BaseBackend._solve_euler/_solve_heun (pyrates/backend/base/base_backend.py:712-769) and the
// NumPy kernels behind the generated vector field: numpy.dot (base_funcs.py:141-142), einsum-wsum
// (:85-88), numpy.sum (:139), numpy.roll ring buffers (:143; pyrates/ir/circuit.py:400-425).
//
// Memory: the state lives in HBM/L2 as flat vectors (Y = y_n, S1..S3 = stage inputs, ACC = RK
// accumulator, YN = y_{n+1}); each element is owned by exactly one thread/warp per evaluation, so only
// the stage vectors are ever read by other CTAs.  W is streamed once per evaluation with 128-bit
// non-allocating loads; the source vector stays L1/L2 resident.
//
// Expects from the generated prologue: real, PRB_NS, PRB_NOBS, PRB_SOLVER, PRB_BLOCK, PRB_NRINGS,
// PRB_WORK_USER (element offset of the solver buffers inside `work`), PRB_RANK, PRB_WORLD, prb_net_init_heads(),
// prb_net_advance_heads(), prb_eval<STAGE>() and PRB_OBS_OFF.

#ifndef PRB_NB
#define PRB_NB 1                 // sweep instances integrated together (batch index b is the fastest-varying dimension)
#endif
#ifndef PRB_LL_REDUNDANT
#define PRB_LL_REDUNDANT 0
#endif
#ifndef PRB_DOT_DEEP
#define PRB_DOT_DEEP 0           // fp64 row dot: 8 instead of 4 independent 16-byte loads per lane (set by the emitter)
#endif

struct PrbNet {
    const real* yin;             // stage input of the current evaluation (written in-kernel: plain loads only)
    const real* __restrict__ params;   // [n_params][PRB_NB] per-instance (swept) scalars
    real* work;
    real* rings;
    const real* __restrict__ consts;
    const int* __restrict__ iconsts;
    const real* __restrict__ tables;
    long long oY, oYN, oS1, oS2, oS3, oACC;     // element offsets of the solver vectors inside `work`
    real dt, hdt, dt6;
    prb_time t;                  // integer step counter (fixed-step solvers) or float time (adaptive RK45)
#if PRB_SOLVER == PRB_SOLVER_RK45
    double h;                    // step size of the current attempt
    double errsq;                // this thread's share of sum((err/scale)^2)
    double rtol, atol;
    long long oK[7], oSA, oSB;   // stage derivatives K0..K6 and the two alternating stage-input vectors inside `work`
#endif
    int head[PRB_NRINGS > 0 ? PRB_NRINGS : 1];
    unsigned long long* bar;     // [0] local arrivals, [1] CTAs done fetching (redundant exchange form), [2] error flag
    unsigned long long bar_target;
    unsigned long long epoch;
#if PRB_WORLD > 1
    uint4* ll_peer[PRB_WORLD];   // receive mirrors of all ranks (IPC-mapped; [PRB_RANK] = this rank's own)
    uint4* ll_mine;
    unsigned long long* peer_sync[PRB_WORLD];
    long long ll_par;            // slot offset of the parity half that the NEXT barrier consumes
    unsigned int ll_tag;         // tag of the packets sent before the next barrier (epoch + 1 with the top bit set: never 0)
#endif
    long long gtid, gsize, gwarp, nwarps;
    int lane;
    unsigned int wcache;         // bit g: the W rows of row group g are cached in shared memory (prb_net_stage_consts)
#ifdef PRB_TC
    PrbTc tc;                    // tcgen05 coupling-GEMM pipeline state (tc_gemm.cuh)
#endif
};

// ---- shared stores: values other CTAs (and, with PRB_WORLD > 1, other GPUs) will read.
// Multi-GPU (one process per GPU, rows of every index space sharded): every rank keeps a complete replica of the work /
// ring arrays.  The owner of an element stores it locally and sends it to every peer as ONE self-validating 16-byte
// packet {lo32, tag, hi32, tag} (the LL protocol: each 8-byte half carries its own flag, and 8-byte stores are single
// transactions on NVLink) into the peer's receive MIRROR — fire-and-forget st.volatile over the IPC mapping, no
// system-scope fence anywhere.  Before a grid barrier the receiving grid polls the mirror slots of the elements its peers
// own (the emitter knows which regions every phase stores: prb_ll_unpack*), writes the values into its local replica and
// then runs the ordinary single-GPU barrier: the data carries the flag, so "the peers have arrived" and "their values are
// here" are the same event.  Mirrors are double-buffered by barrier parity; a per-barrier heartbeat packet from every
// rank keeps all ranks within one barrier of each other, which is what makes two parities enough (a sender can only
// reach the stores of barrier k + 2 after it consumed packets that its peers sent after finishing barrier k).
// r01 (scalar 8-byte replica stores + a system-scope fence + flag round): ~15 us per exchange at 8 GPUs.
#if PRB_WORLD > 1
#define PRB_LL_SLOTS (PRB_LL_WORK + PRB_LL_RING + 32LL)          // work | rings | heartbeats (one slot per rank)
__device__ __forceinline__ void prb_ll_send(const PrbNet& C, const long long slot, const real v) {
    uint4 pk;
    if (sizeof(real) == 8) {
        const unsigned long long u = (unsigned long long)__double_as_longlong((double)v);
        pk.x = (unsigned int)u; pk.z = (unsigned int)(u >> 32);
    } else {
        pk.x = __float_as_uint((float)v); pk.z = 0u;
    }
    pk.y = C.ll_tag; pk.w = C.ll_tag;
#pragma unroll
    for (int p = 0; p < PRB_WORLD; ++p)
        if (p != PRB_RANK)
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};"
                         :: "l"(C.ll_peer[p] + C.ll_par + slot), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
}
#endif
__device__ __forceinline__ void prb_st_work(const PrbNet& C, const long long off, const real v) {
    C.work[off] = v;
#if PRB_WORLD > 1
    prb_ll_send(C, off, v);
#endif
}
__device__ __forceinline__ void prb_st_ring(const PrbNet& C, const long long off, const real v) {
    C.rings[off] = v;
#if PRB_WORLD > 1
    prb_ll_send(C, PRB_LL_WORK + off, v);
#endif
}

// ---- barriers.  Monotonic counter, one arrival per CTA (host zeroes the words before launch).  Spins are bounded: on
// timeout the error word is set (and pushed to the peers) and every later wait falls through, so a lost peer can never
// hang the GPU.
#define PRB_SPIN_LIMIT 20000000000LL

template <bool SYS>
__device__ __forceinline__ unsigned long long prb_ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    // relaxed polling loads; the fence that follows every barrier wait supplies the acquire ordering
    if (SYS) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else     asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void prb_raise_timeout(PrbNet& C) {
    atomicExch(C.bar + 2, 1ULL);
#if PRB_WORLD > 1
#pragma unroll
    for (int p = 0; p < PRB_WORLD; ++p)
        if (p != PRB_RANK) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(C.peer_sync[p] + 2), "l"(1ULL) : "memory");
#endif
}

template <bool SYS>
__device__ __forceinline__ bool prb_spin_until(PrbNet& C, const unsigned long long* p, const unsigned long long target) {
    const long long t0 = clock64();
    unsigned int it = 0;
    while (prb_ld_acquire<SYS>(p) < target) {
        if ((++it & 1023u) == 0) {
            if (prb_ld_acquire<false>(C.bar + 2) != 0ULL) return false;
            if (clock64() - t0 > PRB_SPIN_LIMIT) { prb_raise_timeout(C); return false; }
        }
    }
    return true;
}

#if PRB_WORLD > 1
// wait for the packet in mirror slot `slot` of the current parity and return its payload
__device__ __forceinline__ real prb_ll_recv(PrbNet& C, const long long slot) {
    const uint4* p = C.ll_mine + C.ll_par + slot;
    uint4 pk;
    const long long t0 = clock64();
    unsigned int it = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pk.x), "=r"(pk.y), "=r"(pk.z), "=r"(pk.w) : "l"(p) : "memory");
        if (pk.y == C.ll_tag && pk.w == C.ll_tag) break;
        if ((++it & 1023u) == 0) {
            if (prb_ld_acquire<false>(C.bar + 2) != 0ULL) break;
            if (clock64() - t0 > PRB_SPIN_LIMIT) { prb_raise_timeout(C); break; }
        }
    }
    if (sizeof(real) == 8) return (real)__longlong_as_double((long long)(((unsigned long long)pk.z << 32) | (unsigned long long)pk.x));
    return (real)__uint_as_float(pk.x);
}
// One fetch = one element a peer owns: wait for its packet, write it into the local replica.  The emitter merges all
// regions stored in a phase into ONE grid-strided loop (prb_eval: "fetch what the peers own"), so every thread polls at
// most a few packets whatever the number of regions (12 state variables + a ring column for the QIF-SFA network).
// `q` indexes the elements of the region [0, n) that this rank does NOT own: rows [0, lo) then [lo + own, n).
__device__ __forceinline__ void prb_ll_fetch(PrbNet& C, const long long off, const long long q, const long long lo, const long long own) {
    const long long e = off + (q < lo ? q : q + own);
    C.work[e] = prb_ll_recv(C, e);
}
__device__ __forceinline__ void prb_ll_fetch_ring(PrbNet& C, const long long off, const long long q, const long long lo, const long long own) {
    const long long e = off + (q < lo ? q : q + own);
    C.rings[e] = prb_ll_recv(C, PRB_LL_WORK + e);
}
// scatter stores arr[idx[i]] = v_i, i sharded like every index space: the (constant) index array names the slots
__device__ __forceinline__ void prb_ll_fetch_scatter(PrbNet& C, const long long off, const int* __restrict__ idx, const long long nb,
                                                     const long long q, const long long lo, const long long own) {
    const long long r = q / nb, b = q - r * nb;
    const long long i = r < lo ? r : r + own;
    const long long e = (off + (long long)__ldg(idx + i)) * nb + b;
    C.work[e] = prb_ll_recv(C, e);
}
#endif

// The barrier in two halves.  arrive: this CTA's stores are done (multi-GPU: heartbeat out, peers' heartbeats in).
// wait: all local CTAs have arrived.  prb_grid_barrier = arrive + wait.  In the REDUNDANT exchange form (PRB_LL_REDUNDANT,
// small exchanges) the generated code puts its fetch loop BETWEEN the halves: every CTA polls all peer-owned elements itself
// while the local arrivals trickle in, so the local barrier overlaps the NVLink flight instead of following it.  Word [1] of
// the barrier block counts CTAs that have FINISHED fetching (one increment per CTA and barrier): the heartbeat of barrier
// k + 1 is only sent once every local CTA is done with the packets of barrier k — the lockstep argument that makes two mirror
// parities sufficient needs "this rank has consumed barrier k" to hold for all its CTAs, not just for the arrivals.
__device__ __forceinline__ void prb_barrier_arrive(PrbNet& C) {
#if PRB_WORLD > 1
    if (C.gtid == 0) {
#if PRB_LL_REDUNDANT
        prb_spin_until<false>(C, C.bar + 1, (unsigned long long)gridDim.x * C.epoch);
#endif
        prb_ll_send(C, PRB_LL_WORK + PRB_LL_RING + PRB_RANK, (real)0);
    }
    if (C.gtid >= 32 && C.gtid < 32 + PRB_WORLD && C.gtid - 32 != PRB_RANK) (void)prb_ll_recv(C, PRB_LL_WORK + PRB_LL_RING + (C.gtid - 32));
#endif
    __syncthreads();
    if (threadIdx.x == 0) {
        C.bar_target += gridDim.x;
        __threadfence();
        atomicAdd(C.bar, 1ULL);
    }
}

__device__ __forceinline__ void prb_barrier_wait(PrbNet& C) {
#if PRB_WORLD > 1 && PRB_LL_REDUNDANT
    __syncthreads();                                     // every thread of this CTA is done with the mirror
#endif
    if (threadIdx.x == 0) {
#if PRB_WORLD > 1 && PRB_LL_REDUNDANT
        atomicAdd(C.bar + 1, 1ULL);
#endif
        prb_spin_until<false>(C, C.bar, C.bar_target);
        __threadfence();
    }
    C.epoch += 1ULL;
#if PRB_WORLD > 1
    C.ll_tag = (unsigned int)(C.epoch + 1ULL) | 0x80000000u;
    C.ll_par = (long long)((C.epoch + 1ULL) & 1ULL) * PRB_LL_SLOTS;
#endif
    __syncthreads();
}

__device__ __forceinline__ void prb_grid_barrier(PrbNet& C) {
    prb_barrier_arrive(C);
    prb_barrier_wait(C);
}

__device__ __forceinline__ real prb_warp_sum(real v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result broadcast to every thread (used for full reductions computed redundantly per CTA)
__device__ __forceinline__ real prb_block_sum(real v) {
    __shared__ real red[32];
    __shared__ real total;
    v = prb_warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        real s = (l < (int)(blockDim.x >> 5)) ? red[l] : (real)0;
        s = prb_warp_sum(s);
        if (l == 0) total = s;
    }
    __syncthreads();
    return total;
}

__device__ __forceinline__ double prb_block_sum_d(double v) {
    __shared__ double redd[32];
    __shared__ double totald;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) redd[w] = v;
    __syncthreads();
    if (w == 0) {
        double s = (l < (int)(blockDim.x >> 5)) ? redd[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (l == 0) totald = s;
    }
    __syncthreads();
    return totald;
}

__device__ __forceinline__ long long prb_ring_pos(int head, int col, int L) {
    int p = col - head;
    p += (p < 0) ? L : 0;
    p += (p < 0) ? L : 0;
    return (long long)p;
}

// numpy.interp(x, xp, fp) for one x (timed inputs of adaptive runs, frontend/template/circuit.py:1692-1709): clamp outside
// [xp[0], xp[n-1]], else slope * (x - xp[j]) + fp[j] with xp[j] <= x < xp[j+1] (the expression NumPy's C loop evaluates);
// `stride` > 1 walks one column of a (steps, n_cols) table (interp_rows)
__device__ __forceinline__ real prb_interp(const real* __restrict__ xp, const real* __restrict__ fp, const int stride, const int n,
                                           const real x) {
    if (!(x > __ldg(xp))) return __ldg(fp);
    if (x >= __ldg(xp + n - 1)) return __ldg(fp + (long long)(n - 1) * stride);
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(xp + mid) <= x) lo = mid; else hi = mid;
    }
    const real x0 = __ldg(xp + lo), f0 = __ldg(fp + (long long)lo * stride);
    const real slope = (__ldg(fp + (long long)(lo + 1) * stride) - f0) / (__ldg(xp + lo + 1) - x0);
    return slope * (x - x0) + f0;
}

__device__ __forceinline__ real prb_max(real a, real b) { return (a != a || b != b) ? a + b : (a > b ? a : b); }
__device__ __forceinline__ real prb_min(real a, real b) { return (a != a || b != b) ? a + b : (a < b ? a : b); }
__device__ __forceinline__ real prb_sign(real a) { return (real)((a > (real)0) - (a < (real)0)); }

// ---- streaming row dot product: sum_j w[j] * x[j], lanes stride the row with 128-bit loads.
// w is read-only and touched once per evaluation -> non-coherent, no L1 allocation; x was written by other
// CTAs in the previous evaluation -> coherent loads (L1 was invalidated by the barrier's fence).
__device__ __forceinline__ double2 prb_ld_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 prb_ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ double prb_row_dot(const double* __restrict__ w, const double* x, int m, int lane) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if ((m & 1) == 0 && ((((unsigned long long)w) | ((unsigned long long)x)) & 15ULL) == 0) {
        const double2* w2 = (const double2*)w;
        const double2* x2 = (const double2*)x;
        const int m2 = m >> 1;
        int j = lane;
#if PRB_DOT_DEEP
        // 8 independent 16-byte loads per lane in flight: rows handled by few warps (row-sharded shards, small networks) are
        // bound by the number of bytes in flight, not by the stream rate
        for (; j + 224 < m2; j += 256) {
            double2 wv[8], xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) wv[u] = prb_ld_stream(w2 + j + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = x2[j + 32 * u];
            a0 = fma(wv[0].x, xv[0].x, a0); a0 = fma(wv[0].y, xv[0].y, a0);
            a1 = fma(wv[1].x, xv[1].x, a1); a1 = fma(wv[1].y, xv[1].y, a1);
            a2 = fma(wv[2].x, xv[2].x, a2); a2 = fma(wv[2].y, xv[2].y, a2);
            a3 = fma(wv[3].x, xv[3].x, a3); a3 = fma(wv[3].y, xv[3].y, a3);
            a0 = fma(wv[4].x, xv[4].x, a0); a0 = fma(wv[4].y, xv[4].y, a0);
            a1 = fma(wv[5].x, xv[5].x, a1); a1 = fma(wv[5].y, xv[5].y, a1);
            a2 = fma(wv[6].x, xv[6].x, a2); a2 = fma(wv[6].y, xv[6].y, a2);
            a3 = fma(wv[7].x, xv[7].x, a3); a3 = fma(wv[7].y, xv[7].y, a3);
        }
#endif
        for (; j + 96 < m2; j += 128) {
            const double2 wa = prb_ld_stream(w2 + j), wb = prb_ld_stream(w2 + j + 32);
            const double2 wc = prb_ld_stream(w2 + j + 64), wd = prb_ld_stream(w2 + j + 96);
            const double2 xa = x2[j], xb = x2[j + 32], xc = x2[j + 64], xd = x2[j + 96];
            a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0);
            a1 = fma(wb.x, xb.x, a1); a1 = fma(wb.y, xb.y, a1);
            a2 = fma(wc.x, xc.x, a2); a2 = fma(wc.y, xc.y, a2);
            a3 = fma(wd.x, xd.x, a3); a3 = fma(wd.y, xd.y, a3);
        }
        for (; j < m2; j += 32) {
            const double2 wa = prb_ld_stream(w2 + j);
            const double2 xa = x2[j];
            a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0);
        }
    } else {
        for (int j = lane; j < m; j += 32) a0 = fma(__ldg(w + j), x[j], a0);
    }
    return prb_warp_sum((a0 + a1) + (a2 + a3));
}

__device__ __forceinline__ float prb_row_dot(const float* __restrict__ w, const float* x, int m, int lane) {
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if ((m & 3) == 0 && ((((unsigned long long)w) | ((unsigned long long)x)) & 15ULL) == 0) {
        const float4* w4 = (const float4*)w;
        const float4* x4 = (const float4*)x;
        const int m4 = m >> 2;
        int j = lane;
        for (; j + 32 < m4; j += 64) {
            const float4 wa = prb_ld_stream(w4 + j), wb = prb_ld_stream(w4 + j + 32);
            const float4 xa = x4[j], xb = x4[j + 32];
            a0 = fmaf(wa.x, xa.x, a0); a1 = fmaf(wa.y, xa.y, a1); a2 = fmaf(wa.z, xa.z, a2); a3 = fmaf(wa.w, xa.w, a3);
            a0 = fmaf(wb.x, xb.x, a0); a1 = fmaf(wb.y, xb.y, a1); a2 = fmaf(wb.z, xb.z, a2); a3 = fmaf(wb.w, xb.w, a3);
        }
        for (; j < m4; j += 32) {
            const float4 wa = prb_ld_stream(w4 + j);
            const float4 xa = x4[j];
            a0 = fmaf(wa.x, xa.x, a0); a1 = fmaf(wa.y, xa.y, a1); a2 = fmaf(wa.z, xa.z, a2); a3 = fmaf(wa.w, xa.w, a3);
        }
    } else {
        for (int j = lane; j < m; j += 32) a0 = fmaf(__ldg(w + j), x[j], a0);
    }
    return prb_warp_sum((a0 + a1) + (a2 + a3));
}

// the same dot products with the W row segment in SHARED memory (W-stationary small networks: rows are copied once per
// launch, prb_net_stage_consts); x comes from global / L1 as before
__device__ __forceinline__ double prb_row_dot_sm(const double* w, const double* x, int m, int lane) {
    double a0 = 0, a1 = 0;
    if ((m & 1) == 0 && ((((unsigned long long)w) | ((unsigned long long)x)) & 15ULL) == 0) {
        const double2* w2 = (const double2*)w;
        const double2* x2 = (const double2*)x;
        const int m2 = m >> 1;
        int j = lane;
        for (; j + 32 < m2; j += 64) {
            const double2 wa = w2[j], wb = w2[j + 32], xa = x2[j], xb = x2[j + 32];
            a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0);
            a1 = fma(wb.x, xb.x, a1); a1 = fma(wb.y, xb.y, a1);
        }
        for (; j < m2; j += 32) { const double2 wa = w2[j], xa = x2[j]; a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0); }
    } else {
        for (int j = lane; j < m; j += 32) a0 = fma(w[j], x[j], a0);
    }
    return prb_warp_sum(a0 + a1);
}
__device__ __forceinline__ void prb_row_dot2_sm(const double* w, const double* x, const double* z, int m, int lane, double& rx, double& rz) {
    double a0 = 0, b0 = 0;
    if ((m & 1) == 0 && ((((unsigned long long)w) | ((unsigned long long)x) | ((unsigned long long)z)) & 15ULL) == 0) {
        const double2* w2 = (const double2*)w;
        const double2* x2 = (const double2*)x;
        const double2* z2 = (const double2*)z;
        for (int j = lane; j < (m >> 1); j += 32) {
            const double2 wa = w2[j], xa = x2[j], za = z2[j];
            a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0); b0 = fma(wa.x, za.x, b0); b0 = fma(wa.y, za.y, b0);
        }
    } else {
        for (int j = lane; j < m; j += 32) { const double wv = w[j]; a0 = fma(wv, x[j], a0); b0 = fma(wv, z[j], b0); }
    }
    rx = prb_warp_sum(a0);
    rz = prb_warp_sum(b0);
}
__device__ __forceinline__ float prb_row_dot_sm(const float* w, const float* x, int m, int lane) {
    float a0 = 0;
    for (int j = lane; j < m; j += 32) a0 = fmaf(w[j], x[j], a0);
    return prb_warp_sum(a0);
}
__device__ __forceinline__ void prb_row_dot2_sm(const float* w, const float* x, const float* z, int m, int lane, float& rx, float& rz) {
    float a0 = 0, b0 = 0;
    for (int j = lane; j < m; j += 32) { const float wv = w[j]; a0 = fmaf(wv, x[j], a0); b0 = fmaf(wv, z[j], b0); }
    rx = prb_warp_sum(a0);
    rz = prb_warp_sum(b0);
}

// two right-hand sides over ONE pass of the row (separable pair-space couplings: W.sin(theta) and W.cos(theta))
__device__ __forceinline__ void prb_row_dot2(const double* __restrict__ w, const double* x, const double* z, int m, int lane,
                                             double& rx, double& rz) {
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
    if ((m & 1) == 0 && ((((unsigned long long)w) | ((unsigned long long)x) | ((unsigned long long)z)) & 15ULL) == 0) {
        const double2* w2 = (const double2*)w;
        const double2* x2 = (const double2*)x;
        const double2* z2 = (const double2*)z;
        const int m2 = m >> 1;
        int j = lane;
        for (; j + 32 < m2; j += 64) {
            const double2 wa = prb_ld_stream(w2 + j), wb = prb_ld_stream(w2 + j + 32);
            const double2 xa = x2[j], xb = x2[j + 32], za = z2[j], zb = z2[j + 32];
            a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0); b0 = fma(wa.x, za.x, b0); b0 = fma(wa.y, za.y, b0);
            a1 = fma(wb.x, xb.x, a1); a1 = fma(wb.y, xb.y, a1); b1 = fma(wb.x, zb.x, b1); b1 = fma(wb.y, zb.y, b1);
        }
        for (; j < m2; j += 32) {
            const double2 wa = prb_ld_stream(w2 + j);
            const double2 xa = x2[j], za = z2[j];
            a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0); b0 = fma(wa.x, za.x, b0); b0 = fma(wa.y, za.y, b0);
        }
    } else {
        for (int j = lane; j < m; j += 32) { const double wv = __ldg(w + j); a0 = fma(wv, x[j], a0); b0 = fma(wv, z[j], b0); }
    }
    rx = prb_warp_sum(a0 + a1);
    rz = prb_warp_sum(b0 + b1);
}
__device__ __forceinline__ void prb_row_dot2(const float* __restrict__ w, const float* x, const float* z, int m, int lane,
                                             float& rx, float& rz) {
    float a0 = 0, b0 = 0;
    for (int j = lane; j < m; j += 32) { const float wv = __ldg(w + j); a0 = fmaf(wv, x[j], a0); b0 = fmaf(wv, z[j], b0); }
    rx = prb_warp_sum(a0);
    rz = prb_warp_sum(b0);
}

// ---- batched rows: acc[r][bb] = sum_j w[r][j] * x[j][b0+bb] for R rows and BT sweep instances per warp task.
// The connectivity is shared by all instances of a sweep, so one pass over a W row serves BT instances (and one load
// of x[j][b0..b0+BT) serves R rows): the GEMM-shaped reuse of a grid_search over a network, in fp64 FMA.
template <int R, int BT>
__device__ __forceinline__ void prb_rows_dot_batch(const real* __restrict__ w, const long long ldw, const int rows,
                                                   const real* x, const int m, const int lane, real (&acc)[R][BT]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int bb = 0; bb < BT; ++bb) acc[r][bb] = (real)0;
    for (int j = lane; j < m; j += 32) {
        real xv[BT];
        const real* xp = x + (long long)j * PRB_NB;
#pragma unroll
        for (int bb = 0; bb < BT; ++bb) xv[bb] = xp[bb];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const real wv = (r < rows) ? __ldg(w + r * ldw + j) : (real)0;
#pragma unroll
            for (int bb = 0; bb < BT; ++bb) acc[r][bb] = fma(wv, xv[bb], acc[r][bb]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int bb = 0; bb < BT; ++bb) acc[r][bb] = prb_warp_sum(acc[r][bb]);
}

// lane l of a batched row task owns (row r = l / BT, instance bb = l % BT): pick its sum out of the register tile
template <int R, int BT>
__device__ __forceinline__ real prb_tile_pick(const real (&acc)[R][BT], const int lane) {
    real v = (real)0;
#pragma unroll
    for (int q = 0; q < R * BT; ++q)
        if (lane == q) v = acc[q / BT][q % BT];
    return v;
}

// ---- sweep-batched dense coupling on the FMA pipes (fp64 models, and fp32 batches the tcgen05 path does not take):
// R[i][b] = sum_j W[i][j] * x[j][b] for all sweep instances b that share the connectivity — a GEMM [n x m] x [m x NB].
// CTA tile = 128 rows x TN instances (TN = 32 * CB), K slabs of 16 double-buffered in shared memory (global -> registers
// -> smem prefetch of slab s+1 overlaps the FMAs of slab s, one __syncthreads per slab).  16 warps as 4 (rows) x 4
// (instances); each thread owns an 8 x CB register tile whose rows / instances are INTERLEAVED across the lanes
// (row = 4 r + lane/8, instance = 8 c + lane%8): every shared-memory read is a scalar load of 4 (8) consecutive words
// broadcast inside ONE wavefront — ncu showed the contiguous-tile version bound by 128-bit LDS wavefronts
// (4 per quarter-warp, 537 M bank conflicts) at 42 % of the fp64 pipe.
// Results go to the reduction's work array [i][NB]; the element-wise tail runs as a thread-per-(element, instance)
// phase behind a grid barrier.  Needs blockDim.x == 512.
#define PRB_GT_M 128
#define PRB_GT_K 16
#define PRB_GT_PAD (sizeof(real) == 8 ? 1 : 2)      // As row padding: conflict-free transposing stores
// fp64 variant on the tensor cores: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) runs at the full fp64 rate on B200 (measured
// 36.8 TFLOP/s vs 34.1 for a pure DFMA loop, scripts/dmma_probe.py) with 8x fewer math instructions and 6x fewer
// shared-memory reads per FMA than 8 x 4 register tiles.  (tcgen05 has no fp64 kind: this is the only tensor path for
// the parity precision.)  Warp tile 32 rows x 8*CB instances = 4 x CB fragments; smem: A [row][k] (row stride 20) and
// x [k][inst] (row stride TN + 4) make every fragment load conflict-free.
#define PRB_GD_LDA 20
template <int CB>
__device__ __forceinline__ void prb_gemm_batch_dmma(const PrbNet& C, const double* __restrict__ W, const long long row0, const int n_rows,
                                                    const int m, const double* x, const long long out_off) {
    constexpr int TN = 32 * CB, LDX = TN + 4;
    extern __shared__ __align__(16) unsigned char prb_dyn_smem[];
    double (*As)[PRB_GT_M][PRB_GD_LDA] = (double (*)[PRB_GT_M][PRB_GD_LDA])prb_dyn_smem;                                     // [2][row][k]
    double (*Xs)[PRB_GT_K][LDX] = (double (*)[PRB_GT_K][LDX])(prb_dyn_smem + 2 * PRB_GT_M * PRB_GD_LDA * sizeof(double));   // [2][k][inst]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;                      // fragment coordinates: group id, thread in group
    const int wr0 = (warp >> 2) * 32, wc0 = (warp & 3) * (TN / 4);
    const int a_row = tid >> 2, a_k = (tid & 3) * 4;
    const int x_k = tid >> 5;
    const int n_mb = (n_rows + PRB_GT_M - 1) / PRB_GT_M, n_nb = PRB_NB / TN, KS = (m + PRB_GT_K - 1) / PRB_GT_K;
    const bool w_vec = (m & 3) == 0 && (((unsigned long long)W) & 15ULL) == 0;
    for (int tile = blockIdx.x; tile < n_mb * n_nb; tile += gridDim.x) {
        const int mb = tile / n_nb, nbk = tile - mb * n_nb;
        const int i0 = mb * PRB_GT_M, b0 = nbk * TN;
        double acc[4][CB][2];
#pragma unroll
        for (int fr = 0; fr < 4; ++fr)
#pragma unroll
            for (int fc = 0; fc < CB; ++fc) acc[fr][fc][0] = acc[fr][fc][1] = 0.0;
        double pa[4], px[CB];
        auto load_global = [&](const int s) {
            const int k0 = s * PRB_GT_K;
            const bool row_ok = (i0 + a_row) < n_rows;
            const double* wp = W + (long long)(i0 + a_row) * m + k0 + a_k;
            if (row_ok && w_vec && k0 + a_k + 3 < m) {
                const double2 u = prb_ld_stream((const double2*)wp), v = prb_ld_stream((const double2*)wp + 1);
                pa[0] = u.x; pa[1] = u.y; pa[2] = v.x; pa[3] = v.y;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) pa[c] = (row_ok && k0 + a_k + c < m) ? __ldg(wp + c) : 0.0;
            }
            const bool k_ok = (k0 + x_k) < m;
            const double* xp = x + (long long)(k0 + x_k) * PRB_NB + b0 + lane;
#pragma unroll
            for (int c = 0; c < CB; ++c) px[c] = k_ok ? xp[32 * c] : 0.0;
        };
        auto store_smem = [&](const int buf) {
            *(double2*)&As[buf][a_row][a_k] = make_double2(pa[0], pa[1]);
            *(double2*)&As[buf][a_row][a_k + 2] = make_double2(pa[2], pa[3]);
#pragma unroll
            for (int c = 0; c < CB; ++c) Xs[buf][x_k][lane + 32 * c] = px[c];
        };
        load_global(0);
        __syncthreads();
        store_smem(0);
        __syncthreads();
        for (int s = 0; s < KS; ++s) {
            const int buf = s & 1;
            if (s + 1 < KS) load_global(s + 1);
#pragma unroll
            for (int kq = 0; kq < PRB_GT_K / 4; ++kq) {
                double a[4], bq[CB];
#pragma unroll
                for (int fr = 0; fr < 4; ++fr) a[fr] = As[buf][wr0 + 8 * fr + g][4 * kq + t4];
#pragma unroll
                for (int fc = 0; fc < CB; ++fc) bq[fc] = Xs[buf][4 * kq + t4][wc0 + 8 * fc + g];
#pragma unroll
                for (int fr = 0; fr < 4; ++fr)
#pragma unroll
                    for (int fc = 0; fc < CB; ++fc)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                     : "+d"(acc[fr][fc][0]), "+d"(acc[fr][fc][1]) : "d"(a[fr]), "d"(bq[fc]));
            }
            if (s + 1 < KS) store_smem(buf ^ 1);
            __syncthreads();
        }
#pragma unroll
        for (int fr = 0; fr < 4; ++fr) {
            const int ir = i0 + wr0 + 8 * fr + g;
            if (ir < n_rows) {
#pragma unroll
                for (int fc = 0; fc < CB; ++fc) {
                    const long long o = (out_off + row0 + ir) * PRB_NB + b0 + wc0 + 8 * fc + 2 * t4;
                    prb_st_work(C, o, acc[fr][fc][0]);
                    prb_st_work(C, o + 1, acc[fr][fc][1]);
                }
            }
        }
    }
}

template <int CB>
__device__ __forceinline__ void prb_gemm_batch(const PrbNet& C, const real* __restrict__ W, const long long row0, const int n_rows,
                                               const int m, const real* x, const long long out_off) {
    constexpr int TN = 32 * CB;
    constexpr int LDA = PRB_GT_M + PRB_GT_PAD;
    if constexpr (sizeof(real) == 8) { prb_gemm_batch_dmma<CB>(C, (const double*)W, row0, n_rows, m, (const double*)x, out_off); return; }
    extern __shared__ __align__(16) unsigned char prb_dyn_smem[];
    real (*As)[PRB_GT_K][LDA] = (real (*)[PRB_GT_K][LDA])prb_dyn_smem;                                    // [2][k][row]
    real (*Xs)[PRB_GT_K][TN] = (real (*)[PRB_GT_K][TN])(prb_dyn_smem + 2 * PRB_GT_K * LDA * sizeof(real));   // [2][k][inst]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ri0 = (warp >> 2) * 32 + (lane >> 3);              // this thread's rows inside the tile: ri0 + 4 r
    const int cb0 = (warp & 3) * (TN / 4) + (lane & 7);          // its instances: cb0 + 8 c
    const int a_row = tid >> 2, a_k = (tid & 3) * 4;             // global -> smem: 4 consecutive k of one W row
    const int x_k = tid >> 5;                                    //                 instances lane, lane + 32, .. of one source
    const int n_mb = (n_rows + PRB_GT_M - 1) / PRB_GT_M, n_nb = PRB_NB / TN, KS = (m + PRB_GT_K - 1) / PRB_GT_K;
    const bool w_vec = (m & 3) == 0 && (((unsigned long long)W) & 15ULL) == 0;
    for (int tile = blockIdx.x; tile < n_mb * n_nb; tile += gridDim.x) {
        const int mb = tile / n_nb, nbk = tile - mb * n_nb;
        const int i0 = mb * PRB_GT_M, b0 = nbk * TN;
        real acc[8][CB];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[r][c] = (real)0;
        real pa[4], px[CB];
        auto load_global = [&](const int s) {
            const int k0 = s * PRB_GT_K;
            const bool row_ok = (i0 + a_row) < n_rows;
            const real* wp = W + (long long)(i0 + a_row) * m + k0 + a_k;
            if (row_ok && w_vec && k0 + a_k + 3 < m) {
                if (sizeof(real) == 8) {
                    const double2 u = prb_ld_stream((const double2*)wp), v = prb_ld_stream((const double2*)wp + 1);
                    pa[0] = (real)u.x; pa[1] = (real)u.y; pa[2] = (real)v.x; pa[3] = (real)v.y;
                } else {
                    const float4 u = prb_ld_stream((const float4*)wp);
                    pa[0] = (real)u.x; pa[1] = (real)u.y; pa[2] = (real)u.z; pa[3] = (real)u.w;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) pa[c] = (row_ok && k0 + a_k + c < m) ? __ldg(wp + c) : (real)0;
            }
            const bool k_ok = (k0 + x_k) < m;
            const real* xp = x + (long long)(k0 + x_k) * PRB_NB + b0 + lane;
#pragma unroll
            for (int c = 0; c < CB; ++c) px[c] = k_ok ? xp[32 * c] : (real)0;
        };
        auto store_smem = [&](const int buf) {
#pragma unroll
            for (int c = 0; c < 4; ++c) As[buf][a_k + c][a_row] = pa[c];
#pragma unroll
            for (int c = 0; c < CB; ++c) Xs[buf][x_k][lane + 32 * c] = px[c];
        };
        load_global(0);
        __syncthreads();                    // previous tile's readers are done with both buffers
        store_smem(0);
        __syncthreads();
        for (int s = 0; s < KS; ++s) {
            const int buf = s & 1;
            if (s + 1 < KS) load_global(s + 1);
#pragma unroll
            for (int kk = 0; kk < PRB_GT_K; ++kk) {
                real a[8], xv[CB];
#pragma unroll
                for (int r = 0; r < 8; ++r) a[r] = As[buf][kk][ri0 + 4 * r];
#pragma unroll
                for (int c = 0; c < CB; ++c) xv[c] = Xs[buf][kk][cb0 + 8 * c];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < CB; ++c) acc[r][c] = fma(a[r], xv[c], acc[r][c]);
            }
            if (s + 1 < KS) store_smem(buf ^ 1);
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ir = i0 + ri0 + 4 * r;
            if (ir < n_rows) {
#pragma unroll
                for (int c = 0; c < CB; ++c) prb_st_work(C, (out_off + row0 + ir) * PRB_NB + b0 + cb0 + 8 * c, acc[r][c]);
            }
        }
    }
}

// ---- solver stage logic: what happens to the derivative k of state element e in stage STAGE.
// HEUN_REF reproduces the reference's aliased `heun` (y += dt/2*(k2+k2), base_backend.py:763-765).
#if PRB_SOLVER == PRB_SOLVER_RK45
// Dormand-Prince tableau (emitted by pyrates_b200/rk45.py) — see the RK45 driver at the end of this file
__device__ constexpr double PRB_RK_C[6] = PRB_RK45_C_INIT;
__device__ constexpr double PRB_RK_A[6][5] = PRB_RK45_A_INIT;
__device__ constexpr double PRB_RK_B[6] = PRB_RK45_B_INIT;
__device__ constexpr double PRB_RK_E[7] = PRB_RK45_E_INIT;
__device__ constexpr double PRB_RK_P[7][4] = PRB_RK45_P_INIT;
#endif

template <int STAGE>
__device__ __forceinline__ void prb_emit_k(PrbNet& C, const long long e, const real k) {
    const real y = C.work[C.oY + e];
#if PRB_SOLVER == PRB_SOLVER_RK45
    // STAGE = E - 1 for evaluation E = 1..6 (K_E = k); STAGE 6 = the initial f(t0, y0) -> K0.  Each evaluation also
    // builds, for ITS element only, the input of the next one: y + h * sum_q A[E+1][q] K_q (E < 5), y_new (E = 5).
    constexpr int E = STAGE + 1;
    if constexpr (STAGE == 6) {
        C.work[C.oK[0] + e] = k;
    } else if constexpr (E < 5) {
        C.work[C.oK[E] + e] = k;
        double d = 0.0;
#pragma unroll
        for (int q = 0; q <= E; ++q) d += (q == E ? (double)k : (double)C.work[C.oK[q] + e]) * PRB_RK_A[E + 1][q];
        prb_st_work(C, ((E & 1) ? C.oSB : C.oSA) + e, (real)((double)y + d * C.h));
    } else if constexpr (E == 5) {
        C.work[C.oK[5] + e] = k;
        double d = 0.0;
#pragma unroll
        for (int q = 0; q <= 5; ++q) d += (q == 5 ? (double)k : (double)C.work[C.oK[q] + e]) * PRB_RK_B[q];
        prb_st_work(C, C.oYN + e, (real)((double)y + C.h * d));            // y_new: input of evaluation 6 and the next state
    } else {
        C.work[C.oK[6] + e] = k;
        double er = 0.0;
#pragma unroll
        for (int q = 0; q <= 6; ++q) er += (q == 6 ? (double)k : (double)C.work[C.oK[q] + e]) * PRB_RK_E[q];
        const double scale = C.atol + fmax(fabs((double)y), fabs((double)C.work[C.oYN + e])) * C.rtol;
        const double r = er * C.h / scale;
        C.errsq += r * r;
    }
#elif PRB_SOLVER == PRB_SOLVER_EULER
    prb_st_work(C, C.oYN + e, y + C.dt * k);
#elif PRB_SOLVER == PRB_SOLVER_HEUN_REF
    if (STAGE == 0) prb_st_work(C, C.oS1 + e, y + C.dt * k);
    else            prb_st_work(C, C.oYN + e, y + C.hdt * (k + k));
#elif PRB_SOLVER == PRB_SOLVER_HEUN
    if (STAGE == 0) { C.work[C.oACC + e] = k; prb_st_work(C, C.oS1 + e, y + C.dt * k); }
    else            prb_st_work(C, C.oYN + e, y + C.hdt * (C.work[C.oACC + e] + k));
#elif PRB_SOLVER == PRB_SOLVER_RK4
    if (STAGE == 0)      { C.work[C.oACC + e] = k;                                   prb_st_work(C, C.oS1 + e, y + C.hdt * k); }
    else if (STAGE == 1) { C.work[C.oACC + e] = C.work[C.oACC + e] + (real)2 * k;    prb_st_work(C, C.oS2 + e, y + C.hdt * k); }
    else if (STAGE == 2) { C.work[C.oACC + e] = C.work[C.oACC + e] + (real)2 * k;    prb_st_work(C, C.oS3 + e, y + C.dt * k); }
    else                 prb_st_work(C, C.oYN + e, y + C.dt6 * (C.work[C.oACC + e] + k));
#endif
}

// where prb_emit_k<STAGE> stores the stage input it builds (the region the peers' packets of this evaluation land in)
template <int STAGE>
__device__ __forceinline__ long long prb_stage_target(const PrbNet& C) {
#if PRB_SOLVER == PRB_SOLVER_HEUN_REF || PRB_SOLVER == PRB_SOLVER_HEUN
    return STAGE == 0 ? C.oS1 : C.oYN;
#elif PRB_SOLVER == PRB_SOLVER_RK4
    return STAGE == 0 ? C.oS1 : (STAGE == 1 ? C.oS2 : (STAGE == 2 ? C.oS3 : C.oYN));
#else
    return C.oYN;
#endif
}

template <int STAGE> __device__ void prb_eval(PrbNet& C);     // generated

template <int STAGE>
__device__ __forceinline__ void prb_stage(PrbNet& C, const long long oyin) {
    C.yin = C.work + oyin;
    prb_eval<STAGE>(C);
    prb_net_advance_heads(C);
#if PRB_WORLD > 1 && PRB_LL_REDUNDANT
    prb_barrier_wait(C);                 // the generated evaluation ended with prb_barrier_arrive + its own fetch loop
#else
    prb_grid_barrier(C);
#endif
}

extern "C" __global__ void __launch_bounds__(PRB_BLOCK, 1)
prb_integrate_net(const __grid_constant__ prb_kernel_args A)
{
    PrbNet C;
    C.work = (real*)A.work;
    C.rings = (real*)A.rings;
    C.consts = (const real*)A.consts;
    C.iconsts = A.iconsts;
    C.tables = (const real*)A.tables;
    const long long nsb = (long long)PRB_NS * PRB_NB;        // every state / work / ring array carries the batch dimension
    C.oY = (long long)PRB_WORK_USER * PRB_NB; C.oYN = C.oY + nsb; C.oS1 = C.oY + 2LL * nsb; C.oS2 = C.oY + 3LL * nsb;
    C.oS3 = C.oY + 4LL * nsb; C.oACC = C.oY + 5LL * nsb;
    C.params = (const real*)A.params;
    C.dt = (real)A.dt; C.hdt = (real)(A.dt / 2); C.dt6 = (real)(A.dt / 6);
    C.bar = (unsigned long long*)A.sync;
    C.bar_target = 0;
    C.epoch = 0;
#if PRB_WORLD > 1
#pragma unroll
    for (int p = 0; p < PRB_WORLD; ++p) {
        C.ll_peer[p] = (uint4*)A.peers[p];
        C.peer_sync[p] = (unsigned long long*)A.peers[PRB_WORLD + p];
    }
    C.ll_mine = C.ll_peer[PRB_RANK];
    C.ll_tag = 1u | 0x80000000u;
    C.ll_par = PRB_LL_SLOTS;
#endif
    C.gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    C.gsize = (long long)gridDim.x * blockDim.x;
    C.lane = threadIdx.x & 31;
    C.gwarp = C.gtid >> 5;
    C.nwarps = C.gsize >> 5;
    prb_net_init_heads(C, A.eval0);
    prb_net_stage_consts(C);
#ifdef PRB_TC
    prb_tc_setup(C.tc, C.bar + 2);
#endif

#if PRB_SOLVER == PRB_SOLVER_RK45
    // ---- adaptive: the reference's solver='scipy' == solve_ivp(method='RK45', first_step=dt, t_eval=times)
    // (base_backend.py:802-812) for a network too large for one thread.  Same algorithm as the instance kernel
    // (inst_kernel.cuh): Dormand-Prince stages, SciPy's controller and dense output, K[0] = the most recent rhs
    // evaluation (the reference's aliased buffer).  The whole grid advances with ONE step size: every evaluation is a
    // persistent-grid phase, the RMS error norm is a grid-wide sum (per-CTA partials, one atomicAdd each, read behind the
    // evaluation's barrier), and every thread replays the same controller arithmetic on the same numbers.
    // Options (A.slots, doubles): [0] rtol [1] atol [2] t_start [3] t_bound [4] first_step [5] sample spacing
    // [6] max_step [7] max attempts, [8..10] error-norm accumulators (rotating), [11] status out, [12] attempts out.
    double* opt = (double*)A.slots;
    C.rtol = opt[0]; C.atol = opt[1];
    const double t_bound = opt[3], te_step = opt[5], max_step = opt[6];
    const long long max_attempts = (long long)opt[7];
    double t = opt[2], h_abs = opt[4];
    C.oSA = C.oY + 2LL * nsb; C.oSB = C.oY + 3LL * nsb;
#pragma unroll
    for (int q = 0; q < 7; ++q) C.oK[q] = C.oY + (4LL + q) * nsb;
    C.errsq = 0.0; C.h = 0.0;
    real* extY = (real*)A.y;
    for (long long e = C.gtid; e < nsb; e += C.gsize) C.work[C.oY + e] = extY[e];
    prb_grid_barrier(C);
    C.t = (prb_time)t;
    prb_stage<6>(C, C.oY);                                            // K0 = f(t0, y0)
    real* out = (real*)A.out;
    long long sample = A.sample0;
    const long long n_samples = A.sample0 + A.n_steps;
    long long attempts = 0;
    unsigned int status = 0;
    while (t != t_bound && status == 0) {
        const double min_step = 10.0 * fabs(nextafter(t, __longlong_as_double(0x7ff0000000000000LL)) - t);
        if (h_abs > max_step) h_abs = max_step;
        else if (h_abs < min_step) h_abs = min_step;
        bool rejected = false;
        double h = 0.0, t_new = t;
        for (;;) {
            if (h_abs < min_step) { status = 1; break; }
            if (++attempts > max_attempts) { status = 2; break; }
            t_new = t + h_abs;
            if (t_new - t_bound > 0.0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            C.h = h;
            const int slot = (int)(attempts % 3);
            if (C.gtid == 0) opt[8 + (slot + 1) % 3] = 0.0;           // the accumulator of the NEXT attempt (last read 2 attempts ago)
            // input of evaluation 1: y + h * A[1][0] * K0
            for (long long e = C.gtid; e < nsb; e += C.gsize)
                prb_st_work(C, C.oSA + e, (real)((double)C.work[C.oY + e] + (double)C.work[C.oK[0] + e] * PRB_RK_A[1][0] * h));
            prb_grid_barrier(C);
            C.t = (prb_time)(t + PRB_RK_C[1] * h); prb_stage<0>(C, C.oSA);
            C.t = (prb_time)(t + PRB_RK_C[2] * h); prb_stage<1>(C, C.oSB);
            C.t = (prb_time)(t + PRB_RK_C[3] * h); prb_stage<2>(C, C.oSA);
            C.t = (prb_time)(t + PRB_RK_C[4] * h); prb_stage<3>(C, C.oSB);
            C.t = (prb_time)(t + PRB_RK_C[5] * h); prb_stage<4>(C, C.oSA);
            C.t = (prb_time)(t + h);
            C.yin = C.work + C.oYN;
            prb_eval<5>(C);                                           // K6 = f(t + h, y_new) + this thread's error share
            {
                const double part = prb_block_sum_d(C.errsq);
                C.errsq = 0.0;
                if (threadIdx.x == 0) atomicAdd(opt + 8 + slot, part);
            }
            prb_grid_barrier(C);
            const double sq = *((volatile double*)(opt + 8 + slot));
            const double error_norm = sqrt(sq) / sqrt((double)nsb);
            // K[0] of the next attempt is the newest evaluation whether or not this one is accepted (aliased buffer)
            { const long long tmp = C.oK[0]; C.oK[0] = C.oK[6]; C.oK[6] = tmp; }
            if (error_norm < 1.0) {
                double factor = (error_norm == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(error_norm, -0.2));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                break;
            }
            h_abs *= fmax(0.2, 0.9 * pow(error_norm, -0.2));
            rejected = true;
            if (!(error_norm <= 1.7976931348623157e308)) {
                // the rejected attempt overflowed: its f(t + h, y_new), now aliased as K0, is not finite and would poison
                // every later attempt (the reference dies here with "step size too small"); recover with the true f(t, y)
                C.t = (prb_time)t;
                prb_stage<6>(C, C.oY);
            }
        }
        if (status != 0) break;
        // dense output at the sample times inside (t, t_new]; after the swap K6 sits at oK[0] and the old K0 at oK[6]
        while (sample < n_samples && (double)(sample - A.sample0) * te_step <= t_new) {
            const double x = ((double)(sample - A.sample0) * te_step - t) / h;
            const double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x;
            for (long long q = C.gtid; q < PRB_NOBS; q += C.gsize) {
                const long long e = PRB_OBS_INDEX(q);
                double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
                for (int s_ = 0; s_ < 7; ++s_) {
                    const double kv = (double)C.work[C.oK[s_ == 0 ? 6 : (s_ == 6 ? 0 : s_)] + e];
                    q0 += kv * PRB_RK_P[s_][0]; q1 += kv * PRB_RK_P[s_][1]; q2 += kv * PRB_RK_P[s_][2]; q3 += kv * PRB_RK_P[s_][3];
                }
                out[sample * (long long)PRB_NOBS + q] = (real)(h * (((q0 * p1 + q1 * p2) + q2 * p3) + q3 * p4) + (double)C.work[C.oY + e]);
            }
            ++sample;
        }
        t = t_new;
        { const long long tmp = C.oY; C.oY = C.oYN; C.oYN = tmp; }
    }
    if (C.gtid == 0) { opt[11] = (double)status; opt[12] = (double)attempts; }
#else
    real* extY = (real*)A.y;
    for (long long e = C.gtid; e < nsb; e += C.gsize) C.work[C.oY + e] = extY[e];   // every rank holds a full replica
    prb_grid_barrier(C);

    real* out = (real*)A.out;
    const long long store_step = A.store_step;
    long long until_store = (store_step - (A.step0 % store_step)) % store_step;
    long long sample = A.sample0;
    C.t = A.step0 + A.t0;

    for (long long i = 0; i < A.n_steps; ++i, ++C.t) {
        if (until_store == 0) {
            for (long long q = C.gtid; q < PRB_NOBS * PRB_NB; q += C.gsize) {
                const long long o = q / PRB_NB, b = q - o * PRB_NB;
                out[sample * (long long)PRB_NOBS * PRB_NB + q] = C.work[C.oY + PRB_OBS_INDEX(o) * PRB_NB + b];
            }
            ++sample;
            until_store = store_step;
        }
        --until_store;
#if PRB_SOLVER == PRB_SOLVER_EULER
        prb_stage<0>(C, C.oY);
#elif PRB_SOLVER == PRB_SOLVER_HEUN_REF || PRB_SOLVER == PRB_SOLVER_HEUN
        prb_stage<0>(C, C.oY);
        prb_stage<1>(C, C.oS1);
#elif PRB_SOLVER == PRB_SOLVER_RK4
        prb_stage<0>(C, C.oY);
        prb_stage<1>(C, C.oS1);
        prb_stage<2>(C, C.oS2);
        prb_stage<3>(C, C.oS3);
#endif
        const long long tmp = C.oY; C.oY = C.oYN; C.oYN = tmp;
    }
#endif   // fixed-step solvers
    for (long long e = C.gtid; e < nsb; e += C.gsize) extY[e] = C.work[C.oY + e];
#ifdef PRB_TC
    prb_tc_teardown(C.tc);
#endif
}

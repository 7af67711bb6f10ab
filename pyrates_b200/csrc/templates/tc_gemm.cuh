// tc_gemm.cuh — hand-written tcgen05 / TMEM / bulk-copy (TMA engine) primitives for sm_100a, used by the
// sweep-batched dense coupling  R[i, b] = sum_j W[i, j] * x[j, b]  (numpy.dot of the reference's generated
// vector field, pyrates/backend/base/base_funcs.py:141-142, evaluated for all instances of a grid_search that
// share the connectivity: the one GEMM-shaped piece of the hot path).
//
// Arithmetic: fp32 models only.  tcgen05.mma kind::tf32 keeps 10 mantissa bits per operand, far from the
// rel-1e-4 trajectory tolerance, so every product is error-compensated ("3xTF32"): with x = hi + lo, hi = x with
// the low 13 mantissa bits cleared, lo = x - hi (exact in fp32),
//     W.x ~= Whi.xhi + Whi.xlo + Wlo.xhi        (dropped term ~2^-22 relative, accumulation in fp32 in TMEM).
//
// Operand layout (both operands K-major, SWIZZLE_NONE "interleaved" canonical layout, see cute/arch/mma_sm100_desc.hpp):
// a tile of ROWS x 32 fp32 (one K slab) is stored as core matrices of 8 rows x 16 bytes (4 fp32), contiguous 128 B each:
//     float offset of element (row r, k) inside a slab tile = ((k/4) * (ROWS/8) + r/8) * 32 + (r%8) * 4 + (k%4)
// so   SBO (next 8-row group) = 128 B,  LBO (next 4-wide K core column) = ROWS * 16 B,
// and one MMA (K = 8 tf32 = 2 core columns) advances the descriptor start address by 2 * LBO.
// Tiles are stored in this layout in HBM, so ONE 1-D bulk copy (cp.async.bulk, SASS UBLKCP) brings a whole slab
// tile into shared memory — no tensor map, no swizzle, no smem bank conflicts on the producer side.
#ifndef PRB_TC_GEMM_CUH_
#define PRB_TC_GEMM_CUH_

#define PRB_TC_BK 32                       // fp32 elements per K slab (128 B per row)
#define PRB_TC_SPIN_LIMIT 2000000000LL     // ~1 s: every wait is bounded, a protocol bug cannot hang the GPU

__device__ __forceinline__ unsigned int prb_smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool prb_elect_one() {
    unsigned int pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier
__device__ __forceinline__ void prb_mbar_init(unsigned int bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void prb_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void prb_mbar_expect_tx(unsigned int bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prb_mbar_arrive(unsigned int bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool prb_mbar_try_wait(unsigned int bar, unsigned int parity) {
    unsigned int ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: on timeout the error word is set (checked by the host) and false is returned
__device__ __forceinline__ bool prb_mbar_wait(unsigned int bar, unsigned int parity, unsigned long long* err,
                                              unsigned long long code) {
    if (prb_mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!prb_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > PRB_TC_SPIN_LIMIT) { atomicExch(err, code); return false; }
    }
    return true;
}

// ---- bulk async copy global -> shared (TMA engine, 1-D), completion on an mbarrier
__device__ __forceinline__ void prb_bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- TMEM
__device__ __forceinline__ void prb_tmem_alloc(unsigned int smem_dst, unsigned int ncols) {     // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void prb_tmem_dealloc(unsigned int taddr, unsigned int ncols) {      // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void prb_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void prb_tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 8 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void prb_tmem_ld8(unsigned int taddr, float (&v)[8]) {
    unsigned int r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// ---- descriptors
// shared-memory matrix descriptor (K-major, no swizzle): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48)
__device__ __forceinline__ unsigned long long prb_umma_smem_desc(unsigned int saddr, unsigned int lbo, unsigned int sbo) {
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4)
         | ((unsigned long long)((lbo >> 4) & 0x3FFFu) << 16)
         | ((unsigned long long)((sbo >> 4) & 0x3FFFu) << 32)
         | (1ULL << 46);
}
// instruction descriptor kind::tf32, fp32 accumulate, A and B K-major: c_format=F32 [4,6), a/b_format=TF32 [7,10)/[10,13),
// N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr unsigned int prb_umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned int)(N >> 3) << 17) | ((unsigned int)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void prb_umma_tf32(unsigned int tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                              unsigned int idesc, unsigned int accumulate) {
    asm volatile("{\n\t.reg .pred P;\n\tsetp.ne.b32 P, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, P;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void prb_umma_commit(unsigned int bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// error-compensated split of an fp32 value into two tf32-exact pieces
__device__ __forceinline__ void prb_tf32_split(const float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// float offset of element (row r, k) inside one K-slab tile of ROWS rows (canonical K-major interleaved layout)
__host__ __device__ constexpr long long prb_tc_tile_off(int rows, int r, int k) {
    return ((long long)(k >> 2) * (rows >> 3) + (r >> 3)) * 32 + (r & 7) * 4 + (k & 3);
}
// =====================================================================================================================
// Pipeline used inside the persistent network kernel (net_kernel.cuh) when the emitter defines PRB_TC:
//   PRB_TC_BT      sweep instances per tile (UMMA N; 32 / 64 / 128)
//   PRB_TC_E       epilogue column groups: 4*E epilogue warps (warps 4 .. 4+4E), PRB_TC_BT/E accumulator columns per thread
//   PRB_TC_NSTAGE  shared-memory ring depth (K slabs in flight)
//   PRB_TC_KC      K slabs accumulated inside TMEM before the partial sum is promoted to fp32 registers.  The tensor
//                  core adds into its fp32 accumulator with truncation (measured on B200: K = 8192 gives a relative
//                  bias of -1.7e-4), so partial sums of KC*32 products are added up by the CUDA cores (round-to-nearest).
// Roles: warp 0 lane 0 = bulk-copy producer, warp 1 lane 0 = MMA issuer (warp 1 owns the TMEM allocation),
// warps >= 4 = epilogue (TMEM -> register accumulators -> generated element-wise tail + solver update).
// Two TMEM accumulator buffers of BT columns alternate between the MMA issuer and the epilogue warps.
#ifdef PRB_TC
// Pipeline granularity PRB_TC_CK (K values per shared-memory stage): a K slab of 32 is stored core-column-major, so its first
// 16 K values are the contiguous first half of the tile, in the same canonical layout (same LBO / SBO).  Staging half slabs
// doubles the ring depth inside the same 192 KB: with 256-instance tiles a full-slab stage is 96 KB, i.e. ONE slab in flight
// behind the one being multiplied — the bulk copies' latency was not covered (r01: tensor pipe and its smem reads both 44 %).
#ifndef PRB_TC_CK
#define PRB_TC_CK PRB_TC_BK
#endif
static_assert(PRB_TC_BK % PRB_TC_CK == 0 && PRB_TC_CK % 8 == 0, "stage granularity");
#define PRB_TC_SPLIT (PRB_TC_BK / PRB_TC_CK)
#define PRB_TC_A_BYTES (128 * PRB_TC_CK * 4)
#define PRB_TC_B_BYTES (PRB_TC_BT * PRB_TC_CK * 4)
#define PRB_TC_STAGE_BYTES (2 * PRB_TC_A_BYTES + 2 * PRB_TC_B_BYTES)
#define PRB_TC_SMEM_BYTES (PRB_TC_NSTAGE * PRB_TC_STAGE_BYTES)
#define PRB_TC_NCOLS (2 * PRB_TC_BT)
#define PRB_TC_EPI_WARPS (4 * PRB_TC_E)
#define PRB_TC_CPT (PRB_TC_BT / PRB_TC_E)
static_assert(PRB_TC_NCOLS >= 32 && PRB_TC_NCOLS <= 512 && (PRB_TC_NCOLS & (PRB_TC_NCOLS - 1)) == 0, "TMEM columns: power of two in [32, 512]");
static_assert(PRB_TC_BT % 16 == 0 && PRB_TC_CPT % 8 == 0, "tile shape");

struct PrbTc {
    unsigned int smem0, bar0, tmem;
    int s; unsigned int ph;          // shared-memory ring position (producer thread / MMA thread: one copy each)
    unsigned int buf, uph;           // TMEM accumulator buffer position (MMA thread / epilogue threads)
    unsigned long long* err;
    bool ok;
};
#define PRB_TC_FULL(T, s)   ((T).bar0 + 8u * (s))
#define PRB_TC_EMPTY(T, s)  ((T).bar0 + 8u * (PRB_TC_NSTAGE + (s)))
#define PRB_TC_TFULL(T, b)  ((T).bar0 + 8u * (2 * PRB_TC_NSTAGE + (b)))
#define PRB_TC_TEMPTY(T, b) ((T).bar0 + 8u * (2 * PRB_TC_NSTAGE + 2 + (b)))

// all threads of the CTA, once, at kernel start
__device__ __forceinline__ void prb_tc_setup(PrbTc& T, unsigned long long* err) {
    extern __shared__ __align__(1024) unsigned char prb_tc_smem[];
    __shared__ __align__(8) unsigned long long bars[2 * PRB_TC_NSTAGE + 4];
    __shared__ unsigned int tmem_slot;
    T.smem0 = prb_smem_u32(prb_tc_smem);
    T.bar0 = prb_smem_u32(bars);
    T.s = 0; T.ph = 0; T.buf = 0; T.uph = 0; T.err = err; T.ok = true;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < PRB_TC_NSTAGE; ++s) { prb_mbar_init(PRB_TC_FULL(T, s), 1); prb_mbar_init(PRB_TC_EMPTY(T, s), 1); }
        for (int b = 0; b < 2; ++b) { prb_mbar_init(PRB_TC_TFULL(T, b), 1); prb_mbar_init(PRB_TC_TEMPTY(T, b), PRB_TC_EPI_WARPS); }
        prb_mbar_fence_init();
    }
    if (warp == 1) prb_tmem_alloc(prb_smem_u32(&tmem_slot), PRB_TC_NCOLS);
    prb_tc_fence_before();
    __syncthreads();
    prb_tc_fence_after();
    T.tmem = tmem_slot;
}
// all threads, once, at kernel end
__device__ __forceinline__ void prb_tc_teardown(PrbTc& T) {
    prb_tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) prb_tmem_dealloc(T.tmem, PRB_TC_NCOLS);
}

// generic-proxy global stores -> visible to later bulk copies (async proxy)
__device__ __forceinline__ void prb_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- operand preparation: source vector x[j][NB] (element-major, instance fastest) -> hi / lo K-major tiles
// [NB/BT][KS][BT x 32] in the canonical layout.  One thread handles 4 consecutive j of one instance: coalesced
// reads along the instances, one 16-byte store per piece.  Rows j >= m stay zero (buffers are zero-initialised).
__device__ __forceinline__ void prb_tc_prep(const long long gtid, const long long gsize, const float* x, const int m,
                                            float* xt_hi, float* xt_lo, const int KS) {
    const long long total = (long long)((m + 3) >> 2) * PRB_NB;
    for (long long q = gtid; q < total; q += gsize) {
        const int j4 = (int)(q / PRB_NB), b = (int)(q - (long long)j4 * PRB_NB);
        float v[4], h[4], l[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = 4 * j4 + c;
            v[c] = (j < m) ? x[(long long)j * PRB_NB + b] : 0.0f;
            prb_tf32_split(v[c], h[c], l[c]);
        }
        const int nbk = b / PRB_TC_BT, r = b - nbk * PRB_TC_BT, ks = j4 >> 3, kc = j4 & 7;
        const long long off = ((long long)nbk * KS + ks) * (PRB_TC_BT * PRB_TC_BK) + (kc * (PRB_TC_BT / 8) + (r >> 3)) * 32 + (r & 7) * 4;
        *(float4*)(xt_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
        *(float4*)(xt_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
    prb_fence_proxy_async();
}

// ---- producer (one thread): stream the K slabs of one (row block, instance block) pair through the smem ring.
// a_* / b_* point at the first slab tile; consecutive slabs are one tile apart.
__device__ __forceinline__ void prb_tc_produce(PrbTc& T, const float* a_hi, const float* a_lo, const float* b_hi,
                                               const float* b_lo, const int KS) {
    for (int hs = 0; hs < KS * PRB_TC_SPLIT && T.ok; ++hs) {          // consecutive stages are one (part of a) tile apart
        T.ok = prb_mbar_wait(PRB_TC_EMPTY(T, T.s), T.ph ^ 1u, T.err, 101);
        if (!T.ok) break;
        const unsigned int full = PRB_TC_FULL(T, T.s);
        prb_mbar_expect_tx(full, PRB_TC_STAGE_BYTES);
        const unsigned int base = T.smem0 + T.s * PRB_TC_STAGE_BYTES;
        prb_bulk_g2s(base, a_hi + (long long)hs * (128 * PRB_TC_CK), PRB_TC_A_BYTES, full);
        prb_bulk_g2s(base + PRB_TC_A_BYTES, a_lo + (long long)hs * (128 * PRB_TC_CK), PRB_TC_A_BYTES, full);
        prb_bulk_g2s(base + 2 * PRB_TC_A_BYTES, b_hi + (long long)hs * (PRB_TC_BT * PRB_TC_CK), PRB_TC_B_BYTES, full);
        prb_bulk_g2s(base + 2 * PRB_TC_A_BYTES + PRB_TC_B_BYTES, b_lo + (long long)hs * (PRB_TC_BT * PRB_TC_CK), PRB_TC_B_BYTES, full);
        if (++T.s == PRB_TC_NSTAGE) { T.s = 0; T.ph ^= 1u; }
    }
}

// ---- MMA issuer (one thread): KS slabs, promoted every PRB_TC_KC slabs through alternating TMEM buffers
__device__ __forceinline__ void prb_tc_mma(PrbTc& T, const int KS) {
    constexpr unsigned int idesc = prb_umma_idesc_tf32(128, PRB_TC_BT);
    constexpr unsigned int lboA = 128 * 16, lboB = PRB_TC_BT * 16, sbo = 128;
    for (int ks0 = 0; ks0 < KS && T.ok; ks0 += PRB_TC_KC) {
        T.ok = prb_mbar_wait(PRB_TC_TEMPTY(T, T.buf), T.uph ^ 1u, T.err, 104);
        if (!T.ok) break;
        prb_tc_fence_after();
        const unsigned int tmem_d = T.tmem + T.buf * PRB_TC_BT;
        const int ks1 = (ks0 + PRB_TC_KC < KS) ? ks0 + PRB_TC_KC : KS;
        for (int hs = ks0 * PRB_TC_SPLIT; hs < ks1 * PRB_TC_SPLIT; ++hs) {
            T.ok = prb_mbar_wait(PRB_TC_FULL(T, T.s), T.ph, T.err, 102);
            if (!T.ok) break;
            prb_tc_fence_after();
            const unsigned int base = T.smem0 + T.s * PRB_TC_STAGE_BYTES;
            const unsigned int sa_hi = base, sa_lo = base + PRB_TC_A_BYTES;
            const unsigned int sb_hi = base + 2 * PRB_TC_A_BYTES, sb_lo = sb_hi + PRB_TC_B_BYTES;
#pragma unroll
            for (int k = 0; k < PRB_TC_CK / 8; ++k) {
                const unsigned long long dah = prb_umma_smem_desc(sa_hi + k * 2 * lboA, lboA, sbo);
                const unsigned long long dal = prb_umma_smem_desc(sa_lo + k * 2 * lboA, lboA, sbo);
                const unsigned long long dbh = prb_umma_smem_desc(sb_hi + k * 2 * lboB, lboB, sbo);
                const unsigned long long dbl = prb_umma_smem_desc(sb_lo + k * 2 * lboB, lboB, sbo);
                prb_umma_tf32(tmem_d, dal, dbh, idesc, (hs != ks0 * PRB_TC_SPLIT || k != 0) ? 1u : 0u);     // small terms first
                prb_umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                prb_umma_tf32(tmem_d, dah, dbh, idesc, 1u);
            }
            prb_umma_commit(PRB_TC_EMPTY(T, T.s));          // smem slot free once these MMAs have read it
            if (++T.s == PRB_TC_NSTAGE) { T.s = 0; T.ph ^= 1u; }
        }
        prb_umma_commit(PRB_TC_TFULL(T, T.buf));            // partial sum complete -> epilogue
        T.buf ^= 1u; if (T.buf == 0) T.uph ^= 1u;
    }
}

// ---- epilogue warps: add the partial sums of one reduction into this thread's PRB_TC_CPT accumulator columns
// (row = TMEM lane = 32*q + lane, columns cg*CPT .. +CPT of the instance block)
__device__ __forceinline__ void prb_tc_accumulate(PrbTc& T, const int KS, float (&acc)[PRB_TC_CPT], const int q, const int cg) {
    for (int ks0 = 0; ks0 < KS; ks0 += PRB_TC_KC) {
        bool ok = T.ok && prb_mbar_wait(PRB_TC_TFULL(T, T.buf), T.uph, T.err, 103);
        ok = __all_sync(0xffffffffu, ok);
        T.ok = ok;
        if (!ok) return;
        prb_tc_fence_after();
        const unsigned int taddr = T.tmem + ((unsigned int)(q * 32) << 16) + T.buf * PRB_TC_BT + cg * PRB_TC_CPT;
#pragma unroll
        for (int c = 0; c < PRB_TC_CPT; c += 8) {
            float v[8];
            prb_tmem_ld8(taddr + c, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[c + e] += v[e];
        }
        prb_tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) prb_mbar_arrive(PRB_TC_TEMPTY(T, T.buf));
        T.buf ^= 1u; if (T.buf == 0) T.uph ^= 1u;
    }
}
#endif  // PRB_TC
#endif

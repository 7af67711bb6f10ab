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
#endif

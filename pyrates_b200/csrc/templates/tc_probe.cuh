// tc_probe.cuh — stand-alone check of the tcgen05 building blocks in tc_gemm.cuh: one CTA computes one
// 128 x PRB_TC_BT tile  D = A . B^T  (A: [128 x K], B: [BT x K], both pre-tiled K-major in global memory) with
// the same warp roles the network kernel uses: warp 0 = bulk-copy producer, warp 1 = MMA issuer (+ TMEM
// owner), warps 2..5 = epilogue (TMEM -> registers -> global).  Descriptor fields arrive through `iconsts`
// so the host can sweep them:  [0] K slabs, [1] LBO_A, [2] SBO_A, [3] LBO_B, [4] SBO_B, [5] idesc,
// [6] terms (1 = plain tf32, 3 = 3xTF32), [7] descriptor advance per MMA for A (bytes), [8] same for B.
//   consts = A tiles: hi [MB][KS][128*32] then lo;  work = B tiles: hi [KS][BT*32] then lo;  out = D [MB*128][BT].
#ifndef PRB_TC_BT
#define PRB_TC_BT 128
#endif
#ifndef PRB_TC_NSTAGE
#define PRB_TC_NSTAGE 3
#endif
#define PRB_TC_A_BYTES (128 * PRB_TC_BK * 4)
#define PRB_TC_B_BYTES (PRB_TC_BT * PRB_TC_BK * 4)
#define PRB_TC_STAGE_BYTES (2 * PRB_TC_A_BYTES + 2 * PRB_TC_B_BYTES)
#define PRB_TC_NCOLS (PRB_TC_BT < 32 ? 32 : PRB_TC_BT)

extern "C" __global__ void __launch_bounds__(192, 1)
prb_tc_probe(const __grid_constant__ prb_kernel_args A)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * PRB_TC_NSTAGE + 1];
    __shared__ unsigned int tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KS = A.iconsts[0];
    const unsigned int lboA = A.iconsts[1], sboA = A.iconsts[2], lboB = A.iconsts[3], sboB = A.iconsts[4];
    const unsigned int idesc = (unsigned int)A.iconsts[5];
    const int terms = A.iconsts[6];
    const unsigned int advA = A.iconsts[7], advB = A.iconsts[8];
    const int MB = gridDim.x, mb = blockIdx.x;
    unsigned long long* err = (unsigned long long*)A.sync + 2;

    const unsigned int smem0 = prb_smem_u32(smem);
    const unsigned int bar0 = prb_smem_u32(bars);
    #define FULL(s)  (bar0 + 8u * (s))
    #define EMPTY(s) (bar0 + 8u * (PRB_TC_NSTAGE + (s)))
    #define TFULL    (bar0 + 8u * (2 * PRB_TC_NSTAGE))

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < PRB_TC_NSTAGE; ++s) { prb_mbar_init(FULL(s), 1); prb_mbar_init(EMPTY(s), 1); }
        prb_mbar_init(TFULL, 1);
        prb_mbar_fence_init();
    }
    if (warp == 1) prb_tmem_alloc(prb_smem_u32(&tmem_base_s), PRB_TC_NCOLS);
    prb_tc_fence_before();
    __syncthreads();
    prb_tc_fence_after();
    const unsigned int tmem = tmem_base_s;

    const float* a_hi = (const float*)A.consts;
    const float* a_lo = a_hi + (long long)MB * KS * 128 * PRB_TC_BK;
    const float* b_hi = (const float*)A.work;
    const float* b_lo = b_hi + (long long)KS * PRB_TC_BT * PRB_TC_BK;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; unsigned int ph = 0;
            for (int ks = 0; ks < KS; ++ks) {
                if (!prb_mbar_wait(EMPTY(s), ph ^ 1u, err, 101)) break;
                prb_mbar_expect_tx(FULL(s), PRB_TC_STAGE_BYTES);
                const unsigned int base = smem0 + s * PRB_TC_STAGE_BYTES;
                const long long aoff = ((long long)mb * KS + ks) * 128 * PRB_TC_BK;
                const long long boff = (long long)ks * PRB_TC_BT * PRB_TC_BK;
                prb_bulk_g2s(base, a_hi + aoff, PRB_TC_A_BYTES, FULL(s));
                prb_bulk_g2s(base + PRB_TC_A_BYTES, a_lo + aoff, PRB_TC_A_BYTES, FULL(s));
                prb_bulk_g2s(base + 2 * PRB_TC_A_BYTES, b_hi + boff, PRB_TC_B_BYTES, FULL(s));
                prb_bulk_g2s(base + 2 * PRB_TC_A_BYTES + PRB_TC_B_BYTES, b_lo + boff, PRB_TC_B_BYTES, FULL(s));
                if (++s == PRB_TC_NSTAGE) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int s = 0; unsigned int ph = 0;
            bool ok = true;
            for (int ks = 0; ks < KS && ok; ++ks) {
                ok = prb_mbar_wait(FULL(s), ph, err, 102);
                if (!ok) break;
                prb_tc_fence_after();
                const unsigned int base = smem0 + s * PRB_TC_STAGE_BYTES;
                const unsigned int sa_hi = base, sa_lo = base + PRB_TC_A_BYTES;
                const unsigned int sb_hi = base + 2 * PRB_TC_A_BYTES, sb_lo = sb_hi + PRB_TC_B_BYTES;
#pragma unroll
                for (int k = 0; k < PRB_TC_BK / 8; ++k) {
                    const unsigned long long dah = prb_umma_smem_desc(sa_hi + k * advA, lboA, sboA);
                    const unsigned long long dbh = prb_umma_smem_desc(sb_hi + k * advB, lboB, sboB);
                    if (terms == 3) {
                        const unsigned long long dal = prb_umma_smem_desc(sa_lo + k * advA, lboA, sboA);
                        const unsigned long long dbl = prb_umma_smem_desc(sb_lo + k * advB, lboB, sboB);
                        prb_umma_tf32(tmem, dal, dbh, idesc, (ks | k) != 0);
                        prb_umma_tf32(tmem, dah, dbl, idesc, 1u);
                        prb_umma_tf32(tmem, dah, dbh, idesc, 1u);
                    } else {
                        prb_umma_tf32(tmem, dah, dbh, idesc, (ks | k) != 0);
                    }
                }
                prb_umma_commit(EMPTY(s));
                if (++s == PRB_TC_NSTAGE) { s = 0; ph ^= 1u; }
            }
            prb_umma_commit(TFULL);
        }
    } else {
        if (prb_mbar_wait(TFULL, 0u, err, 103)) {
            prb_tc_fence_after();
            const int q = warp & 3;
            float* out = (float*)A.out + ((long long)mb * 128 + q * 32 + lane) * PRB_TC_BT;
#pragma unroll 1
            for (int c = 0; c < PRB_TC_BT; c += 8) {
                float v[8];
                prb_tmem_ld8(tmem + ((unsigned int)(q * 32) << 16) + c, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) out[c + e] = v[e];
            }
        }
    }
    prb_tc_fence_before();
    __syncthreads();
    if (warp == 1) prb_tmem_dealloc(tmem, PRB_TC_NCOLS);
}

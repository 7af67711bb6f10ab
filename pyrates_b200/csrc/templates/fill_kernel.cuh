// fill_kernel.cuh — connectivity-sized constants generated ON the device (pyrates_b200/devgen.py).
//
// The C5 configuration (Kuramoto, N = 65536) carries a dense 32 GiB fp64 weight matrix: drawing it with NumPy and
// pushing it over PCIe would take minutes of a benchmark that integrates for a fraction of a second, and under
// row sharding every rank only needs its own rows.  prb_fill_uniform writes dst[e] = scale * u01(seed, first + e) with a
// counter-based generator (SplitMix64 finaliser of seed + (index + 1) * golden ratio): every element depends only on its
// global index, so any row shard can be generated independently and the NumPy twin (DeviceUniform.host) reproduces the
// matrix bit for bit for the parity tests.  Arguments ride in prb_kernel_args: y = dst, n_inst = element count,
// step0 = first global element index, t0 = seed, dt = scale.  `real` is typedef'ed by the generated prologue.
__device__ __forceinline__ unsigned long long prb_splitmix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

extern "C" __global__ void __launch_bounds__(256)
prb_fill_uniform(const __grid_constant__ prb_kernel_args A)
{
    real* dst = (real*)A.y;
    const unsigned long long seed = (unsigned long long)A.t0, first = (unsigned long long)A.step0;
    const long long n = A.n_inst, stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const unsigned long long z = prb_splitmix64(seed + (first + (unsigned long long)e + 1ULL) * 0x9E3779B97F4A7C15ULL);
        dst[e] = (real)(A.dt * ((double)(z >> 11) * 0x1.0p-53));
    }
}

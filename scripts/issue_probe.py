"""B200 micro-benchmark: does an FP64-pipe instruction leave the issue slot of its second pipe cycle to other pipes?

Per iteration every thread issues 32 independent DFMAs plus n integer / fp32 instructions of a given kind (independent
chains, `asm volatile` so that the mix survives the compiler).  16 warps per SM (4 per scheduler, the C2 kernel's occupancy).
If the schedulers overlap, the time stays that of the 32 DFMAs (64 pipe cycles) until n > 32; if a DFMA blocks its
scheduler for both cycles, the time grows like 64 + n.
    python scripts/issue_probe.py            (GPU)      DUMP=1 python scripts/issue_probe.py   (CPU: SASS mix only)
"""
import json, os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrates_b200 import runtime as rt
from pyrates_b200.emit_inst import read_template

KINDS = {
    "iadd": 'x[{j}] += x[{j5}];',
    "lop": 'x[{j}] ^= x[{j5}];',
    "imad": 'x[{j}] = x[{j}] * x[{j5}] + k;',
    "ffma": 'f[{j}] = fmaf(f[{j}], f[{j5}], fk);',
    "lds": 'asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(x[{j}]) : "r"(sa + ((x[{j}] & 511u) << 2)));',
    "mufu": 'asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[{j}]));',
    "i2d": 'g[{j}] = (double)(int)x[{j}]; x[{j}] += k;',
    # second round: the other instruction kinds of the C2 loop
    "vimnmx": 'x[{j}] = max(max(x[{j}], x[{j5}]), x[{j7}]);',
    "ldsb": 'asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(x[{j}]) : "r"(sa + ((x[{j}] & 0x1fcu))));',   # few distinct words per warp
    "ldc": 'asm volatile("ld.const.f64 %0, [%1];" : "=d"(g[{j}]) : "l"(cb + ((it + {j}) & 15)));',
    "shf": 'x[{j}] = __funnelshift_l(x[{j5}], x[{j}], 3);',
    "branch": 'if (x[{j}] == 0xdeadbeefu) {{ x[{j}] = probe_rare(x[{j}]); }} x[{j}] += k;',
    "vote": 'if (__any_sync(0xffffffffu, x[{j}] == 0xdeadbeefu)) {{ x[{j}] = probe_rare(x[{j}]); }} x[{j}] += k;',
    "mov64": 'asm volatile("mov.b64 %0, 0x3fc5555500000000;" : "=d"(g[{j}]));',
}


def kernel(name, kind, n, ndf=32, warps_note=""):
    body = []
    per = [n * (i + 1) // ndf - n * i // ndf for i in range(ndf)]        # spread the n extras between the DFMAs
    j = 0
    for i in range(ndf):
        if kind == "dfma3":            # three distinct register operands per DFMA (register-file read model: 3 cycles each)
            body.append(f'asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[{i}]) : "d"(c[{(i + 7) % 32}]), "d"(c[{(i + 13) % 32}]));')
            continue
        body.append(f'asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[{i}]) : "d"(a), "d"({'g[%d]' % (i % 16) if kind in ('i2d', 'ldc', 'mov64') else 'b'}));')
        for _ in range(per[i]):
            body.append(KINDS[kind].format(j=j % 16, j5=(j + 5) % 16, j7=(j + 7) % 16))
            j += 1
    return f'''
extern "C" __global__ void __launch_bounds__(512, 1) {name}(const __grid_constant__ prb_kernel_args A) {{
    __shared__ unsigned int sm[512];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const unsigned int sa = (unsigned int)__cvta_generic_to_shared(sm);
    const double* cb = probe_cb;
    double c[32], g[16];
    unsigned int x[16]; float f[16];
    for (int i = 0; i < 32; ++i) c[i] = threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 16; ++i) {{ x[i] = threadIdx.x + i; f[i] = threadIdx.x * 1e-3f + i; g[i] = 0.0; }}
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    const unsigned int k = (unsigned int)A.n_inst | 1u; const float fk = 1.0f + 1e-7f * (float)A.n_inst;
    for (long long it = 0; it < A.n_steps; ++it) {{
        {chr(10).join("        " + l for l in body)}
    }}
    double s = 0; for (int i = 0; i < 32; ++i) s += c[i];
    for (int i = 0; i < 16; ++i) s += (double)x[i] + (double)f[i] + g[i];
    ((double*)A.out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}}
'''


PRE = r"""
__device__ __constant__ double probe_cb[32] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32};
__device__ __noinline__ unsigned int probe_rare(unsigned int v) { return v * 2654435761u + 12345u; }
"""
CASES2 = [("base", "iadd", 0)] + [(f"{k}{n}", k, n) for k in ("vimnmx", "ldsb", "ldc", "shf", "mov64") for n in (8, 16)] + \
         [("branch4", "branch", 4), ("branch8", "branch", 8), ("vote4", "vote", 4), ("vote8", "vote", 8), ("imad64", "imad", 64)]
CASES = [("base", "iadd", 0)] + [(f"{k}{n}", k, n) for k in ("iadd", "lop", "imad", "ffma") for n in (16, 32)] + \
        [("iadd64", "iadd", 64), ("lds8", "lds", 8), ("lds16", "lds", 16), ("mufu8", "mufu", 8), ("mufu16", "mufu", 16),
         ("i2d8", "i2d", 8), ("i2d16", "i2d", 16), ("dfma3", "dfma3", 0)]

if __name__ == "__main__":
    if os.environ.get("ROUND") == "2":
        CASES = CASES2
    src = read_template("prb_kernel_abi.cuh") + PRE + "".join(kernel("probe_" + name, kind, n) for name, kind, n in CASES)
    mod = rt.compile_module(src, "issue_probe.cu")
    if os.environ.get("DUMP"):
        with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as fh:
            fh.write(mod.cubin)
        for name, kind, n in CASES:
            sass = subprocess.run(["cuobjdump", "-sass", "-fun", "probe_" + name, fh.name], capture_output=True, text=True).stdout
            ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4})\*/\s+(.*?);", sass)]
            loops = [(int(re.search(r"0x([0-9a-f]+)", s).group(1), 16), a) for a, s in ins if "BRA" in s and re.search(r"0x([0-9a-f]+)", s)]
            lo, hi = max((l for l in loops if l[0] < l[1]), key=lambda l: l[1] - l[0])
            ops = [s.split()[0] if not s.startswith("@") else s.split()[1] for a, s in ins if lo <= a <= hi]
            import collections
            print(name, collections.Counter(o.split(".")[0] for o in ops).most_common(6))
        sys.exit(0)
    d_out = rt.DeviceBuffer(148 * 512 * 8)
    d_y = rt.DeviceBuffer(64)
    iters = 20000
    res = {}
    for name, kind, n in CASES:
        a = rt.KernelArgs()
        a.n_steps, a.store_step, a.n_inst, a.y, a.out = iters, 1, 1, d_y.ptr, d_out.ptr
        mod.run("probe_" + name, a, block=512, grid=148)
        rt.synchronize()
        e0, e1 = rt.Event(), rt.Event()
        e0.record(); mod.run("probe_" + name, a, block=512, grid=148); e1.record(); e1.synchronize()
        ms = e0.elapsed_ms(e1)
        cyc = ms * 1e-3 * 1.965e9 / iters / 4           # cycles per iteration and warp at 4 warps per scheduler (nominal clock)
        res[name] = {"kind": kind, "extra": n, "ms": round(ms, 4), "cycles_per_warp_iter": round(cyc, 2)}
        print(f"[issue_probe] {name}: 32 DFMA + {n} {kind}: {ms:.3f} ms = {cyc:.1f} cycles per warp-iteration (64 = FP64 pipe alone)")
    print(json.dumps(res))

"""Hardware probe (run on the GPU box): which shared-memory layout / descriptor reading makes a bf16 MN-major
(row-per-sample) SS-mode tcgen05.mma compute D = A^T B?  Prints the max error of every candidate."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "selftest"))
import binding as L

lib = L.lib()
torch.manual_seed(0)
dev = "cuda:0"
for N in (64, 32, 96, 128):
    A = torch.randn(128, 64, device=dev).bfloat16().float()
    B = torch.randn(128, N, device=dev).bfloat16().float()
    ref = (A.double().t() @ B.double()).float()            # [64, N]
    cands = []
    GA, GB = (64 // 8) * 128, (N // 8) * 128
    # fmt 0: interleaved core matrices; try both LBO/SBO readings
    cands += [("interleave lbo=kgroup sbo=128", 0, 0, GA, GB, 128, 2 * GA, 2 * GB),
              ("interleave lbo=128 sbo=kgroup(A)", 0, 0, 128, 128, GA, 2 * GA, 2 * GB),
              ("interleave lbo=128 sbo=kgroup(B)", 0, 0, 128, 128, GB, 2 * GA, 2 * GB)]
    for lt in (4, 5, 3):
        cands += [(f"sw64 type{lt} lbo=8192 sbo=512", 1, lt, 8192, 8192, 512, 1024, 1024),
                  (f"sw64 type{lt} lbo=512 sbo=8192", 1, lt, 512, 512, 8192, 1024, 1024)]
    for lt in (2, 1):
        cands += [(f"sw128 type{lt} lbo=16384 sbo=1024", 2, lt, 16384, 16384, 1024, 2048, 2048),
                  (f"sw128 type{lt} lbo=1024 sbo=16384", 2, lt, 1024, 1024, 16384, 2048, 2048)]
    for name, fmt, lt, lbo_a, lbo_b, sbo, ka, kb in cands:
        D = torch.zeros(128, N, device=dev)
        L.check(lib.dfpt_probe_bf16(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, fmt, lt, lbo_a, lbo_b, sbo, ka, kb, None))
        torch.cuda.synchronize()
        rows = torch.tensor([(m // 16) * 32 + m % 16 for m in range(64)], device=dev)
        got = D[rows]
        err = (got - ref).abs().max().item()
        print(f"N={N:4d} {name:40s} max_err={err:.3e} {'OK' if err < 1e-3 else ''}")

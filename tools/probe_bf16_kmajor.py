"""Hardware probe for the attention chain (run on the GPU box): the plane layout read as a K-major bf16 A operand with the
sample as the MN dimension, and the forward weight image read transposed (MN-major B) for the backward chain.  Prints
the max error of D for one candidate descriptor reading per invocation (a wrong reading can fault, so every candidate
runs in its own process):   for i in 0 1 2 3 4 5; do python tools/probe_bf16_kmajor.py $i; done"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "selftest"))
import torch
import binding as L

lib = L.lib()
torch.manual_seed(0)
dev = "cuda:0"
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if which < 4:
    for K, N in ((64, 128), (128, 128), (128, 64), (32, 32)):
        A = torch.randn(128, K, device=dev).bfloat16().float()
        W = torch.randn(N, K, device=dev).bfloat16().float()
        ref = (A.double() @ W.double().t()).float()
        cands = [("A: lbo=128 sbo=16K | W: lbo=N*16 sbo=128", 128, 16 * K, N * 16, 128),
                 ("A: lbo=128 sbo=16K | W: lbo=128 sbo=N*16", 128, 16 * K, 128, N * 16),
                 ("A: lbo=16K sbo=128 | W: lbo=N*16 sbo=128", 16 * K, 128, N * 16, 128),
                 ("A: lbo=16K sbo=128 | W: lbo=128 sbo=N*16", 16 * K, 128, 128, N * 16)]
        name, la, sa, lb, sb = cands[which]
        D = torch.zeros(128, N, device=dev)
        L.check(lib.dfpt_probe_bf16_k(A.data_ptr(), W.data_ptr(), D.data_ptr(), K, N, la, sa, lb, sb, None))
        torch.cuda.synchronize()
        err = (D - ref).abs().max().item()
        print(f"K={K:4d} N={N:4d} {name:44s} max_err={err:.3e} {'OK' if err < 1e-3 else ''}", flush=True)
else:
    print("---- transposed read of the forward weight image (backward chain): D = A[128,J] @ Wt[J,KIN]")
    for J, KIN in ((64, 128), (128, 128), (128, 64), (32, 32)):
        A = torch.randn(128, J, device=dev).bfloat16().float()
        Wt = torch.randn(J, KIN, device=dev).bfloat16().float()
        ref = (A.double() @ Wt.double()).float()
        cands = [("A: lbo=128 sbo=16J | Wt: lbo=128 sbo=J*16", 128, 16 * J, 128, J * 16),
                 ("A: lbo=128 sbo=16J | Wt: lbo=J*16 sbo=128", 128, 16 * J, J * 16, 128)]
        name, la, sa, lb, sb = cands[which - 4]
        D = torch.zeros(128, KIN, device=dev)
        L.check(lib.dfpt_probe_bf16_kt(A.data_ptr(), Wt.data_ptr(), D.data_ptr(), J, KIN, la, sa, lb, sb, None))
        torch.cuda.synchronize()
        err = (D - ref).abs().max().item()
        print(f"J={J:4d} KIN={KIN:4d} {name:44s} max_err={err:.3e} {'OK' if err < 1e-3 else ''}", flush=True)

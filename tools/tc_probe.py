"""Decode the tcgen05 node GEMM's operand mapping with integer-valued probes (exact in TF32)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshdqn_b200 import _lib
L, p = _lib.lib(), _lib.ptr
dev = torch.device("cuda:0")

def run(A, W, mode=1):
    M, K = A.shape
    kpad = (K + 7) // 8 * 8
    Ap = torch.zeros(M, kpad); Ap[:, :K] = A
    wp = torch.zeros(kpad, 128); wp[:K] = W
    hi = (wp.view(torch.int32) & -8192).view(torch.float32)
    tile = lambda m: m.view(kpad // 4, 4, 16, 8).permute(0, 2, 3, 1).contiguous().view(-1)
    ws = torch.cat([tile(hi), tile(wp - hi)]).to(dev)
    C = torch.full((M, 128), -7.0, device=dev)
    Ad, Wd = Ap.to(dev), W.contiguous().to(dev)      # keep the device tensors alive across the launch
    _lib.check(L.mdq_node_gemm(p(Ad), None, kpad, K, M, 128, p(Wd), p(ws), None, None, None, 0, mode, p(C), None, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return C.cpu()

torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
for K in (8, 16, 32):
    g = torch.Generator().manual_seed(K)
    A = torch.randint(-3, 4, (128, K), generator=g).float()
    W = torch.randint(-3, 4, (K, 128), generator=g).float()
    C = run(A, W)
    ref = A @ W
    print("K", K, "random int: max err", float((C - ref).abs().max()), "fp32 path err", float((run(A, W, 0) - ref).abs().max()))
    # probes
    for kk in range(min(K, 8)):
        A = torch.zeros(128, K); A[:, kk] = torch.arange(128).float()
        W = torch.zeros(K, 128); W[kk, :] = 1.0
        C = run(A, W)
        ok_rows = bool((C == torch.arange(128).float()[:, None]).all())
        A = torch.zeros(128, K); A[:, kk] = 1.0
        W = torch.zeros(K, 128); W[kk, :] = torch.arange(128).float()
        C2 = run(A, W)
        ok_cols = bool((C2 == torch.arange(128).float()[None, :]).all())
        print("  k", kk, "row-map ok", ok_rows, "col-map ok", ok_cols)
        if not ok_rows: print("   C[:12,0] =", C[:12, 0].tolist(), " C[64:70,0] =", C[64:70, 0].tolist())
        if not ok_cols: print("   C2[0,:12] =", C2[0, :12].tolist(), " C2[0,64:70] =", C2[0, 64:70].tolist())
    # cross-k probe: A has k=a only, W has k=b only -> must be zero unless a == b
    for a in range(min(K, 8)):
        A = torch.zeros(128, K); A[:, a] = 1.0
        hits = []
        for b in range(min(K, 8)):
            W = torch.zeros(K, 128); W[b, :] = 1.0
            C = run(A, W)
            if float(C.abs().max()) != 0: hits.append((b, float(C[0, 0])))
        print("  A k=%d pairs with W k:" % a, hits)

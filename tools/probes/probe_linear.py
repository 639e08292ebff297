"""Stand-alone GPU probe (not collected by pytest): prints per-block error maps for hg_linear."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
abi = importlib.import_module("3dhumangan_b200.abi")
abi.require_device()
torch.manual_seed(0)
for (M, K, N, passes) in [(128, 64, 256, 1), (128, 64, 256, 3), (128, 256, 256, 3), (256, 256, 512, 3)]:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    img, Nb = abi.pack_weight(W)
    Y = abi.linear(X, img, Nb, N, passes=passes); torch.cuda.synchronize()
    ref = (X.double() @ W.double().t()).float()
    err = ((Y - ref).norm() / ref.norm()).item()
    print(f"M={M} K={K} N={N} passes={passes}: rel={err:.3e} finite={bool(torch.isfinite(Y).all())}")
    if err > 1e-2:
        E = (Y - ref).abs().reshape(M // 32, 32, N // 32, 32).amax(dim=(1, 3))
        print("max-abs error per 32x32 block (rows=M/32, cols=N/32):"); print(E.cpu().numpy().round(2))
        # does Y equal some simple mis-mapping?
        Xb = X.bfloat16().float(); Wb = W.bfloat16().float()
        print("vs bf16 ref:", ((Y - Xb @ Wb.t()).norm() / ref.norm()).item())

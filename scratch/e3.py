"""no-swizzle K = 16 bias MMA: which (lbo, sbo) convention is right?  (mode 3: lbo 128 / sbo 256, mode 11: swapped)"""
import ctypes, torch, b200pkg
pkg = b200pkg.load()
DEV = "cuda:0"
torch.manual_seed(0)
A, W, b = torch.randn(128, 64, device=DEV), torch.randn(128, 64, device=DEV), torch.randn(128, device=DEV) * 3
B = torch.cat([W.reshape(-1), b]).contiguous()
want = A.double() @ W.double().T + b.double()[None, :]
for mode in (0, 3, 11):
    C = torch.full((128, 128), float("nan"), device=DEV)
    rc = pkg.LIB.rec_debug_tc_gemm(mode, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(C.data_ptr()),
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    w = want if mode else want - b.double()[None, :]
    print("mode", mode, "rc", rc, "max err", (C.double() - w).abs().max().item(), "scale", w.abs().max().item())

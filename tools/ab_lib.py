"""Same-box A/B of two builds of libhdrnet_b200.so at the headline shape (box-to-box spread is
+-4 %, larger than most kernel changes): python tools/ab_lib.py old.so new.so [variant]
Both libraries are driven directly through the C-ABI (hdrnet_slice_apply_f32_ws), interleaved
bursts, median reported."""
import ctypes, statistics, sys, torch

paths = sys.argv[1:3]
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
B, H, W, GH, GW, GD = 8, 2160, 3840, 16, 16, 8
gen = torch.Generator(device="cuda").manual_seed(1234)
grid = torch.rand(B, GH, GW, GD, 12, device="cuda", generator=gen)
guide = torch.rand(B, H, W, device="cuda", generator=gen)
inp = torch.rand(B, H, W, 3, device="cuda", generator=gen)
outs = [torch.empty_like(inp) for _ in paths]
libs = []
for p in paths:
    lib = ctypes.CDLL(p)
    lib.hdrnet_slice_apply_workspace_bytes.restype = ctypes.c_size_t
    lib.hdrnet_slice_apply_workspace_bytes.argtypes = [ctypes.c_int] * 4
    lib.hdrnet_slice_apply_f32_ws.restype = ctypes.c_int
    lib.hdrnet_slice_apply_f32_ws.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 10 + [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    n = lib.hdrnet_slice_apply_workspace_bytes(B, H, GW, GD)
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    libs.append((lib, ws, n))
stream = torch.cuda.current_stream().cuda_stream

def run(i):
    lib, ws, n = libs[i]
    rc = lib.hdrnet_slice_apply_f32_ws(grid.data_ptr(), guide.data_ptr(), inp.data_ptr(), outs[i].data_ptr(),
                                       B, H, W, GH, GW, GD, 3, 3, 1, variant, ws.data_ptr(), n, stream)
    assert rc == 0, rc

def burst(i, iters=40):
    for _ in range(3): run(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): run(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

res = [[] for _ in paths]
for r in range(9):
    for i in range(len(paths)): res[i].append(burst(i))
for i, p in enumerate(paths):
    med = statistics.median(res[i])
    print(f"{p:48s} median {med:.4f} ms  min {min(res[i]):.4f}  frac {B*H*W*28/med/1e6/6577.4:.4f}")
print("max |A - B| =", float((outs[0] - outs[1]).abs().max()))

import ctypes as C, time, numpy as np, torch
rt = C.CDLL("libcudart.so.12")
n = 10_000_000; pitch = 124
host = np.zeros(n * pitch, np.uint8)
torch.cuda.cudart().cudaHostRegister(host.ctypes.data, host.nbytes, 0)
dev = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
full = t(lambda: rt.cudaMemcpyAsync(dev.data_ptr(), host.ctypes.data, n * pitch, 1, None))
print(f"H2D full 1.24GB: {full*1e3:.1f} ms  {n*pitch/full/1e9:.1f} GB/s")
for off, w in [(0, 40), (68, 56), (44, 80), (0, 124)]:
    h2d = t(lambda: rt.cudaMemcpy2DAsync(dev.data_ptr() + off, pitch, host.ctypes.data + off, pitch, w, n, 1, None))
    d2h = t(lambda: rt.cudaMemcpy2DAsync(host.ctypes.data + off, pitch, dev.data_ptr() + off, pitch, w, n, 2, None))
    print(f"2D width {w:3d}: H2D {h2d*1e3:.1f} ms ({n*w/h2d/1e9:.1f} GB/s)  D2H {d2h*1e3:.1f} ms ({n*w/d2h/1e9:.1f} GB/s)")
d2hf = t(lambda: rt.cudaMemcpyAsync(host.ctypes.data, dev.data_ptr(), n * pitch, 2, None))
print(f"D2H full: {d2hf*1e3:.1f} ms {n*pitch/d2hf/1e9:.1f} GB/s")

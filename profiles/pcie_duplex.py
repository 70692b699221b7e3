"""Host<->device copy floor of the e2e leg: 1.31 GB H2D and 1.31 GB D2H from / to pinned memory, alone and concurrently
(two streams), whole buffers and in 64 chunks per direction."""
import time, torch
n = 1310720000 // 8
h_in = torch.zeros(n, dtype=torch.float64).pin_memory()
h_out = torch.zeros(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = n // chunks
    for k in range(chunks):
        sl = slice(k * c, (k + 1) * c)
        if h2d:
            with torch.cuda.stream(s1):
                d_in[sl].copy_(h_in[sl], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3
for name, a, b, ch in (("H2D alone", 1, 0, 1), ("D2H alone", 0, 1, 1), ("both, whole buffers", 1, 1, 1), ("both, 64 chunks", 1, 1, 64),
                       ("both, 320 chunks (64 x 5 fields)", 1, 1, 320)):
    run(a, b, ch)
    ms = min(run(a, b, ch) for _ in range(5))
    print(f"{name:36s} {ms:7.2f} ms  ({1.31072 / ms * 1e3:.1f} GB/s per direction)")

# the same bytes as 2-D copies (cudaMemcpy2DAsync through ctypes): 16 segments (z-layers) per copy, like the chunked host step
# (5 fields x 64 y-chunks per direction = 320 copies of 16 x 256 KB)
import ctypes, glob, os
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = ctypes.CDLL(cands[0])
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                 ctypes.c_int, ctypes.c_void_p]
L = n // 16
def run2d(chunks):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = L // chunks
    for k in range(chunks):
        off = k * c * 8
        rt.cudaMemcpy2DAsync(d_in.data_ptr() + off, L * 8, h_in.data_ptr() + off, L * 8, c * 8, 16, 1, s1.cuda_stream)
        rt.cudaMemcpy2DAsync(h_out.data_ptr() + off, L * 8, d_out.data_ptr() + off, L * 8, c * 8, 16, 2, s2.cuda_stream)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3
for ch in (64, 320, 640, 1280):
    run2d(ch)
    ms = min(run2d(ch) for _ in range(5))
    print(f"both, {ch:4d} 2-D copies of 16 x {L // ch * 8 // 1024} KB   {ms:7.2f} ms  ({1.31072 / ms * 1e3:.1f} GB/s per direction)")

import time, torch
from cuda import cudart
rows, w, pitch = 22936, 5734, 5760
src = torch.zeros((rows, pitch), dtype=torch.int16, device="cuda")
dst = torch.empty((rows, w), dtype=torch.int16).pin_memory()
dst1 = torch.empty((rows * w,), dtype=torch.int16).pin_memory()
st = torch.cuda.Stream()
def t2d():
    cudart.cudaMemcpy2DAsync(dst.data_ptr(), w * 2, src.data_ptr(), pitch * 2, w * 2, rows, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, st.cuda_stream)
def t1d():
    cudart.cudaMemcpyAsync(dst1.data_ptr(), src.data_ptr(), rows * w * 2, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, st.cuda_stream)
for name, fn in (("2D", t2d), ("1D", t1d), ("2D", t2d), ("1D", t1d)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"D2H {name}: {dt*1e3:.2f} ms  {rows*w*2/dt/1e9:.1f} GB/s")

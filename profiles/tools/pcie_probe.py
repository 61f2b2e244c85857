"""Raw host<->device copy bandwidth of the box (pinned vs write-combined, H2D alone vs with a concurrent D2H), at 1 GPU
or under torchrun at N GPUs: the ceiling of bench.py's e2e leg.  gpurun -- 'python profiles/tools/pcie_probe.py'"""
import ctypes, os, sys, time, torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = ctypes.CDLL("libcudart.so.12")
N = 1 << 30
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(N, dtype=torch.uint8, device="cuda")
def alloc(flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 1, N)
    return p
def bw(src_h, dst_h, both):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), src_h, ctypes.c_size_t(N), 1, ctypes.c_void_p(s1.cuda_stream))
        if both: rt.cudaMemcpyAsync(dst_h, ctypes.c_void_p(dev2.data_ptr()), ctypes.c_size_t(N), 2, ctypes.c_void_p(s2.cuda_stream))
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    return 4 * N / (time.perf_counter() - t0) / 1e9
pin, wc, dst = alloc(0), alloc(4), alloc(0)
for name, h in (("pinned", pin), ("write-combined", wc)):
    a = bw(h, dst, False); b = bw(h, dst, True)
    if rank == 0: print(f"world={world} H2D from {name}: alone {a:.1f} GB/s, with concurrent D2H {b:.1f} GB/s per direction")

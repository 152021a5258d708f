import os, time, torch, torch.distributed as dist
lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
buf = torch.zeros(1925774, device="cuda")
for _ in range(5): dist.all_reduce(buf)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(50): dist.all_reduce(buf)
e[1].record(); torch.cuda.synchronize()
if lr == 0: print("all_reduce 7.7 MB: %.3f ms each" % (e[0].elapsed_time(e[1]) / 50), flush=True)
# all-reduce while a long kernel chain runs on the compute stream (big matmuls)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
def work():
    for _ in range(20): a @ a
work(); torch.cuda.synchronize()
t0 = time.perf_counter(); work(); torch.cuda.synchronize(); tw = time.perf_counter() - t0
t0 = time.perf_counter(); work(); dist.all_reduce(buf); work(); dist.all_reduce(buf); torch.cuda.synchronize(); tb = time.perf_counter() - t0
if lr == 0: print("20 matmuls %.2f ms ; 2x(20 matmuls + all_reduce) %.2f ms" % (tw * 1e3, tb * 1e3), flush=True)
dist.destroy_process_group()

"""Can this box capture NCCL collectives (the ones the sharded step uses) in a CUDA graph?  torchrun, 2+ ranks."""
import os, sys, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def say(*a):
    print(f"[r{rank} {time.strftime('%X')}]", *a, flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "thread_local"
x = torch.ones(1 << 20, device=dev) * (rank + 1)
send = [3 + rank, 5][:world] if world == 2 else [4] * world
cnt = torch.tensor(send, device=dev)
rc = torch.empty_like(cnt)
dist.all_to_all_single(rc, cnt)
recv = rc.tolist()
a2a_in = torch.arange(sum(send), device=dev, dtype=torch.float32) + 100 * rank
a2a_out = torch.empty(sum(recv), device=dev)
ag_out = torch.empty(world * 1024, device=dev)
rs_out = torch.empty(1024, device=dev)
def step():
    y = x * 2
    dist.all_reduce(y)
    dist.all_to_all_single(a2a_out, a2a_in, recv, send)
    dist.all_gather_into_tensor(ag_out, y[:1024].contiguous())
    dist.reduce_scatter_tensor(rs_out, ag_out.clone())
    return y.sum() + a2a_out.sum() + rs_out.sum()
say("eager", step().item())
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
say("warm-up done, capturing with capture_error_mode =", mode)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, capture_error_mode=mode):
    out = step()
say("captured")
for i in range(3):
    g.replay()
    torch.cuda.synchronize()
    say("replay", i, out.item())
dist.barrier()
say("done")
dist.destroy_process_group()

import os, time, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
x=torch.zeros(1,dtype=torch.float64,device="cuda"); h=torch.zeros(1,dtype=torch.float64).pin_memory()
for _ in range(50):
    dist.all_reduce(x, op=dist.ReduceOp.MAX); h.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(1000):
    dist.all_reduce(x, op=dist.ReduceOp.MAX); h.copy_(x, non_blocking=True); torch.cuda.synchronize()
t1=time.perf_counter()
y=torch.zeros(30000*4,dtype=torch.float64,device="cuda"); z=torch.zeros_like(y)
peer=(rank+1)%dist.get_world_size(); prev=(rank-1)%dist.get_world_size()
for _ in range(20):
    ops=[dist.P2POp(dist.isend,y,peer),dist.P2POp(dist.irecv,z,prev)]
    for r in dist.batch_isend_irecv(ops): r.wait()
    torch.cuda.synchronize()
t2=time.perf_counter()
for _ in range(500):
    ops=[dist.P2POp(dist.isend,y,peer),dist.P2POp(dist.irecv,z,prev)]
    for r in dist.batch_isend_irecv(ops): r.wait()
    torch.cuda.synchronize()
t3=time.perf_counter()
if rank==0: print("allreduce+d2h+sync us", (t1-t0)*1e3, " sendrecv 960KB + sync us", (t3-t2)/500*1e6, flush=True)
dist.destroy_process_group()

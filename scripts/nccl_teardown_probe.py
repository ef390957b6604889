"""Why did destroy_process_group() hang after NCCL all-reduces were captured in the step's CUDA graph?  Runs a few
graph-replayed DP steps, then tears down in the mode given by MODE and reports how long the teardown took (a
watchdog thread exits the process after 30 s).
  MODE=plain        destroy_process_group() with the graph still alive
  MODE=drop_graph   delete the step (graph, streams), gc, synchronize, then destroy
  MODE=abort        drop the graph, then ProcessGroupNCCL abort instead of destroy"""
import gc, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
mode = os.environ.get("MODE", "plain")
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
dist.init_process_group("nccl", device_id=dev)
from importlib import import_module
import vaegan_b200 as vb
VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
torch.manual_seed(1)
nets = [vb.Encoder([3, 64, 64], 128), vb.Generator(nz=128, hw=64), vb.Discriminator(hw=64)]
for m in nets: m.to(dev)
step = VAEGANStep(*nets, use_cuda_graph=True)
x = (torch.rand(32, 3, 64, 64) * 2 - 1).to(dev)
for _ in range(5): step.step(x, 50)
torch.cuda.synchronize(); dist.barrier()
def watchdog():
    time.sleep(30); print(f"rank {rank} MODE={mode}: teardown HUNG (>30 s)", flush=True); os._exit(3)
threading.Thread(target=watchdog, daemon=True).start()
t0 = time.time()
if mode in ("drop_graph", "abort"):
    step._graph = None; del step; gc.collect(); torch.cuda.synchronize()
if mode == "abort":
    dist.distributed_c10d._get_default_group()._get_backend(dev).abort()
else:
    dist.destroy_process_group()
print(f"rank {rank} MODE={mode}: teardown took {time.time() - t0:.2f} s", flush=True)

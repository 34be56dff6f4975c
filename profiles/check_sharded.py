"""torchrun --nproc-per-node N profiles/check_sharded.py : the sharded run over N GPUs (NCCL all-gather of results)
reproduces the single-GPU run of the same batch bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from said_b200.model.diffusion import SAID_UNet1D
from said_b200.parallel import sharded_inference
from said_b200.synth import synthetic_batch, synthetic_state_dict

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = SAID_UNet1D()
m.load_state_dict(synthetic_state_dict(0))
m.to(f"cuda:{local}").eval()
B = 2 * world + 1            # ragged on purpose
wave = synthetic_batch(B, 1.0)
with torch.no_grad():
    got = sharded_inference(m, wave, seed=5, num_inference_steps=10, guidance_scale=2.0)
    gen = torch.Generator(device=f"cuda:{local}"); gen.manual_seed(5)
    noise = torch.randn(B, 60, 32, device=f"cuda:{local}", generator=gen)
    full = m._run(wave.to(f"cuda:{local}"), noise, None, None, 10, 1.0, 2.0, 0.0, 0.0, 60, False, False, None).result
ok = torch.equal(got, full)
print(f"rank {rank}/{world}: sharded == single-device: {ok}; shape {tuple(got.shape)}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

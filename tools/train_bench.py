"""Time of one training iteration (Trainer._train_iteration, src/train/training.py:177-207: forward in train() mode,
MSE, backward, Adam step) on the CUDA path, batch = 400 patches (one 320x320 slice) as in the reference's configs.
    python tools/train_bench.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import siren as osiren
from mri_inr_b200.modulated_siren import ModulatedSiren
from mri_inr_b200.tiling import extract_center_batch

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 400
sd = osiren.synth_state_dict(seed=12, mod_bias_shift=0.5)
m = ModulatedSiren(dim_in=2, dim_hidden=256, dim_out=1, num_layers=5, latent_dim=256, w0=1.0, w0_initial=30.0, use_bias=True,
                   dropout=0.1, modulate=True, encoder_type="custom", encoder_path=None, outer_patch_size=32,
                   inner_patch_size=16, siren_patch_size=24, device=torch.device("cpu"), activation="sine")
m.load_state_dict(sd, strict=True)
m.to(DEV).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
under = torch.rand(B, 32, 32, device=DEV)
target = extract_center_batch(torch.rand(B, 32, 32, device=DEV), 32, 24)


def it():
    opt.zero_grad()
    loss = torch.nn.functional.mse_loss(m(under), target)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    it()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record()
out = m(under)
e[1].record()
loss = torch.nn.functional.mse_loss(out, target)
loss.backward()
e[2].record()
torch.cuda.synchronize()
fwd, bwd = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
e[0].record()
n = 10
for _ in range(n):
    it()
e[3].record()
torch.cuda.synchronize()
step = e[0].elapsed_time(e[3]) / n
flops = 3 * B * 576 * 4 * 2 * 256 * 256
print(f"batch {B} patches: forward {fwd:.2f} ms, loss+backward {bwd:.2f} ms, full iteration (zero_grad, forward, backward, "
      f"Adam) {step:.2f} ms = {B / step * 1e3:.0f} patches/s, {flops / step / 1e9:.0f} TFLOP/s of hidden-layer work "
      f"(forward + dX + dW)")

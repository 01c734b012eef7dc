"""Host-side cost of one ComputeBatchLoss + backward at a small batch (the strong-scaling regime of configs[2]:
128 samples per GPU at N = 8): cProfile of 300 steps, top functions by cumulative time, and the step time."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import folax_b200
from folax_b200.loss_functions import ThermalLoss2DQuad

B = int(os.environ.get("B", 128))
mesh = folax_b200.create_2D_square_mesh(1.0, 257)
loss = ThermalLoss2DQuad("t", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}, "beta": 2.0, "c": 4}, mesh)
loss.Initialize()
nn = mesh.GetNumberOfNodes()
g = torch.Generator(device="cuda").manual_seed(0)
K = torch.rand((B, nn), generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
u = torch.rand((B, nn), generator=g, device="cuda", dtype=torch.float64)


def step():
    uu = u.detach().requires_grad_(True)
    kk = K.detach().requires_grad_(True)
    mean, _ = loss.ComputeBatchLoss(kk, uu)
    mean.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step()
t1 = time.perf_counter()       # host time to ENQUEUE 300 steps
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B}: host enqueue {1e3 * (t1 - t0) / 300:.3f} ms/step, wall {1e3 * (t2 - t0) / 300:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])

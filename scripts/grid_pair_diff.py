"""Debug helper: float32 grid kernel with one vs two samples per lane on the same inputs (FOL_ENERGY_GRID_PAIR is read
once per process, so each variant runs in its own process and dumps its gradients)."""
import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from tests.test_zy6_energy_grid_gpu import grid_mesh, make
    mesh = grid_mesh(70, 23)
    loss = make(mesh, dtype="float32")
    rng = np.random.default_rng(62)
    B = 2
    K = rng.uniform(0.1, 1.0, (B, mesh.GetNumberOfNodes())).astype(np.float32)
    u = rng.uniform(0.1, 1.0, (B, mesh.GetNumberOfNodes())).astype(np.float32)
    e, gu, gk = loss._energy_and_grads(torch.tensor(K, device="cuda"), torch.tensor(u, device="cuda"),
                                       dir_values=loss._dir_full, dir_flag=loss._dir_flag, out_scale=1.0)
    np.save(sys.argv[1], np.stack([gu.cpu().numpy(), gk.cpu().numpy()]))
    print(e.cpu().numpy())
else:
    for pair in ("0", "1"):
        subprocess.run([sys.executable, __file__, f"/tmp/grid_pair_{pair}.npy"], env=dict(os.environ, FOL_ENERGY_GRID_PAIR=pair), check=True)
    a, b = np.load("/tmp/grid_pair_0.npy"), np.load("/tmp/grid_pair_1.npy")
    d = np.abs(a - b)
    print("max abs diff gu, gk:", d[0].max(), d[1].max(), "scale", np.abs(a[0]).max(), np.abs(a[1]).max())
    idx = np.argwhere(d[0] > 0)
    print("differing gu entries:", len(idx), "of", a[0].size, "first:", idx[:10].tolist())
    cols = np.unique(idx[:, 1] % 71) if len(idx) else []
    print("columns with differences:", list(cols)[:40])

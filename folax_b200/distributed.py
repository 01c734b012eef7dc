"""Multi-GPU paths of the hot path (one process per GPU, torch.distributed / NCCL plumbing).

* ``SlabPartition`` -- a single huge structured mesh split into contiguous element slabs along z
  (the reference has no domain decomposition; this is the north-star's
  element-partition + halo-DOF exchange).  Each rank owns its slab's elements (so its block of
  the duplicate-keeping BCOO and its Gauss-point state need no exchange) and holds the nodes
  those elements touch; the node planes between two slabs exist on both neighbours.
  ``halo_sum`` adds the neighbour's partial sums on those planes to the local residual, after
  which both copies hold the globally assembled value.
* ``allreduce_gradients`` -- the data-parallel FOL path: samples are sharded over ranks and the
  network gradients are summed with one bucketed NCCL all-reduce (the reference lets XLA insert
  it: deep_network.py:225-242).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import mesh as _mesh


class SlabPartition:
    def __init__(self, Nx, Ny, Nz_global, Lx, Ly, Lz, rank, world, element_type="hexahedron"):
        if Nz_global % world:
            raise ValueError("Nz must be divisible by the number of ranks")
        self.rank, self.world = rank, world
        self.Nx, self.Ny, self.Nz_global = Nx, Ny, Nz_global
        nz = Nz_global // world
        self.nz_local = nz
        hz = Lz / Nz_global
        maker = _mesh.create_3D_box_mesh if element_type == "hexahedron" else _mesh.create_3D_tetra_box_mesh
        m = maker(Nx, Ny, nz, Lx, Ly, hz * nz)
        X = np.array(m.nodes_coordinates)
        X[:, 2] += rank * nz * hz
        m.nodes_coordinates = X
        self.mesh = m
        self.element_type = element_type
        plane = (Nx + 1) * (Ny + 1)
        self.plane_nodes = plane
        self.lower_nodes = np.arange(plane, dtype=np.int64)                  # z = slab bottom
        self.upper_nodes = np.arange(nz * plane, (nz + 1) * plane, dtype=np.int64)
        # local node -> global node (global numbering = the same generator on the whole box)
        self.node_offset = rank * nz * plane
        self.element_offset = rank * m.GetNumberOfElements(element_type)

    def global_node_ids(self):
        return np.arange(self.mesh.GetNumberOfNodes(), dtype=np.int64) + self.node_offset

    def halo_index(self, dofs_per_node, device):
        """Flat dof indices of the lower / upper interface planes (contiguous ranges, because the
        generator numbers nodes plane by plane)."""
        d = dofs_per_node
        lo = (0, self.plane_nodes * d)
        up = (self.nz_local * self.plane_nodes * d, (self.nz_local + 1) * self.plane_nodes * d)
        return lo, up

    def halo_sum(self, residual, dofs_per_node, group=None):
        """In-place neighbour sum of the interface-plane residual partial sums.  The planes are
        contiguous slices of the local residual, so no pack / unpack kernels are needed: the
        slices are sent as they are and the received partials are added."""
        if self.world == 1:
            return residual
        (l0, l1), (u0, u1) = self.halo_index(dofs_per_node, residual.device)
        ops, recv_lo, recv_up = [], None, None
        if self.rank > 0:
            recv_lo = torch.empty(l1 - l0, dtype=residual.dtype, device=residual.device)
            ops.append(dist.P2POp(dist.isend, residual[l0:l1], self.rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, recv_lo, self.rank - 1, group))
        if self.rank < self.world - 1:
            recv_up = torch.empty(u1 - u0, dtype=residual.dtype, device=residual.device)
            ops.append(dist.P2POp(dist.isend, residual[u0:u1], self.rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, recv_up, self.rank + 1, group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if recv_lo is not None:
            residual[l0:l1] += recv_lo
        if recv_up is not None:
            residual[u0:u1] += recv_up
        return residual


    # ---- NVLink peer-memory exchange (csrc/halo.cu): gather of an interface plane fused with the transfer
    _halo = None
    _halo_step = 0

    def enable_peer_halo(self, loss, group=None):
        """Create this rank's receive buffers, swap CUDA IPC handles with the neighbours (one all_gather of 64
        bytes per rank) and open theirs.  After this, assemble_overlapped pushes the interface-plane partial sums
        straight into the neighbour's memory from the gather kernel instead of calling NCCL send/recv."""
        import ctypes as C
        from . import _lib
        if self.world == 1:
            return False
        lib = _lib.load()
        d = loss.number_dofs_per_node
        h = C.c_void_p()
        _lib.check(lib.fol_halo_create(C.byref(h), loss._dt, self.plane_nodes * d))
        raw = (C.c_ubyte * 64)()
        _lib.check(lib.fol_halo_export(h, raw))
        mine = torch.tensor(list(raw), dtype=torch.uint8, device=loss.device)
        handles = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(handles, mine, group=group)
        for side, nb in ((0, self.rank - 1), (1, self.rank + 1)):
            if 0 <= nb < self.world:
                buf = (C.c_ubyte * 64)(*handles[nb].cpu().tolist())
                _lib.check(lib.fol_halo_connect(h, side, buf))
        self._halo, self._halo_step = h, 0
        return True

    def close_peer_halo(self):
        """Releases the peer buffers; raises if any arrival wait gave up (a neighbour never pushed)."""
        if self._halo is not None:
            from . import _lib
            torch.cuda.synchronize()
            lost = _lib.load().fol_halo_timeouts(self._halo)
            _lib.load().fol_halo_destroy(self._halo)
            self._halo = None
            if lost:
                raise _lib.FolaxError(f"halo exchange: {lost} arrival waits timed out (a neighbour rank never pushed)")


FUSED_HALO = True          # False: force the layered path (A/B in tests and scripts/halo_diag.py)
GRID_MARGIN_CTAS = 0   # measured on 2/4/8 B200: a margin does not pay; NCCL's kernels fit next to the persistent CTAs


def assemble_overlapped(loss, part, controls, dofs, ke_out, comm_stream, group=None, state_in=None, state_out=None,
                        kernel_events=None):
    """Residual + Jacobian of one slab with the halo-DOF exchange hidden behind the element stage.

    The element layers touching the slab interfaces are contiguous element ranges (the generator
    numbers elements layer by layer), and the interface planes are contiguous node ranges, so:
      1. element stage on the bottom and top element layers (two small launches),
      2. deterministic residual gather of the two interface node planes,
      3. on `comm_stream`: neighbour exchange + add of those planes (NCCL send/recv over NVLink),
      4. meanwhile on the main stream: element stage on the interior layers, gather of all interior
         nodes,
      5. main stream joins the communication stream.
    History-dependent elements (J2 elastoplasticity, BASELINE.json configs[4]): `state_in` / `state_out` are this
    slab's Gauss-point history (ne, g, 7|4); the state is element-local, so like the Jacobian blocks it needs no
    exchange -- every element-range launch reads and writes its own slice.
    kernel_events: optional (start, end) torch events recorded around the element-stage launch of the fused path
    (bench.py times the dominant kernel inside the measured step with them).
    Returns (ke_out, residual)."""
    from . import _lib
    lib = _lib.load()
    dt, d, A, nd = loss._dt, loss.number_dofs_per_node, loss._nnode, loss._nd
    ne, nn = loss._ne, loss._nn
    esz = 8 if loss.dtype == torch.float64 else 4
    layer = ne // part.nz_local                     # elements per z-layer
    plane = part.plane_nodes
    K = _lib.to_device(controls, loss.dtype).reshape(-1)
    u = _lib.to_device(dofs, loss.dtype).reshape(-1)
    re = torch.empty(ne * nd, dtype=loss.dtype, device=loss.device)
    R = torch.empty(loss.total_number_of_dofs, dtype=loss.dtype, device=loss.device)
    phys, elem = _lib.PHYSICS[loss.physics], loss.fe_element.code
    s = _lib.stream_ptr()
    st_bytes = 0                                    # bytes of Gauss-point history per element
    if state_in is not None or state_out is not None:
        if loss.physics != "j2plasticity":
            raise ValueError("assemble_overlapped: Gauss-point state is an input of the elastoplastic losses only")
        if state_in is None or state_out is None:
            raise ValueError("assemble_overlapped: state_in and state_out go together")
        state_in = _lib.to_device(state_in, loss.dtype)
        if (not state_out.is_cuda or not state_out.is_contiguous() or state_out.dtype != loss.dtype
                or state_out.numel() != state_in.numel() or state_in.numel() % max(ne, 1)):
            raise ValueError("assemble_overlapped: state_out must be a contiguous CUDA tensor shaped like state_in "
                             "(ne, g, 7|4)")
        st_bytes = esz * (state_in.numel() // max(ne, 1))

    def elements(e0, cnt):
        if cnt <= 0:
            return
        _lib.check(lib.fol_assemble_elements(s, dt, phys, elem, loss.num_gp, 0, cnt, nn, _lib.ptr(loss._xyz),
                                             loss._conn.data_ptr() + 4 * A * e0, _lib.ptr(K), _lib.ptr(u),
                                             _lib.ptr(loss._dir_flag), loss._params,
                                             ke_out.data_ptr() + esz * nd * nd * e0, re.data_ptr() + esz * nd * e0,
                                             (state_in.data_ptr() + st_bytes * e0) if st_bytes else None,
                                             (state_out.data_ptr() + st_bytes * e0) if st_bytes else None))

    def gather(n0, cnt):
        if cnt <= 0:
            return
        _lib.check(lib.fol_residual_gather(s, dt, cnt, A, d, loss._adj_ptr.data_ptr() + 4 * n0, _lib.ptr(loss._adj),
                                           _lib.ptr(re), R.data_ptr() + esz * d * n0))

    if part.world > 1 and part._halo is not None:
        step, h = part._halo_step, part._halo
        part._halo_step += 1
        if (FUSED_HALO and dt == _lib.F64 and elem == 0 and loss.num_gp == 2 and d == 3
                and loss.physics in ("mechanical", "j2plasticity")):
            # fused path: ONE element-stage launch that visits the interface layers first and pushes the two planes
            # over NVLink from inside (csrc/assemble_hex_common.cuh), ONE launch for the interior gather + halo add
            if kernel_events is not None:
                kernel_events[0].record()
            _lib.check(lib.fol_assemble_elements_halo(
                s, dt, phys, elem, loss.num_gp, ne, nn, _lib.ptr(loss._xyz), _lib.ptr(loss._conn), _lib.ptr(K),
                _lib.ptr(u), _lib.ptr(loss._dir_flag), loss._params, _lib.ptr(ke_out), _lib.ptr(re),
                _lib.ptr(state_in) if st_bytes else None, _lib.ptr(state_out) if st_bytes else None,
                h, step, layer, plane, _lib.ptr(loss._adj_ptr), _lib.ptr(loss._adj), _lib.ptr(R)))
            if kernel_events is not None:
                kernel_events[1].record()
            _lib.check(lib.fol_residual_gather_halo(s, h, step, nn, plane, _lib.ptr(loss._adj_ptr), _lib.ptr(loss._adj),
                                                    _lib.ptr(re), _lib.ptr(R)))
            return ke_out, R
        # layered NVLink peer path: plane gather + push in one kernel, add after the interior work (no NCCL, no side
        # stream); serves the element types / precisions the fused kernels do not cover

        def push(side, n0):
            _lib.check(lib.fol_halo_gather_push(s, h, side, step, n0, plane, d, _lib.ptr(loss._adj_ptr),
                                                _lib.ptr(loss._adj), _lib.ptr(re), _lib.ptr(R)))
        if part.nz_local == 1:
            elements(0, ne)
        else:
            elements(0, layer)
            elements(ne - layer, layer)
        push(0, 0)
        push(1, nn - plane)
        if part.nz_local > 2:
            elements(layer, ne - 2 * layer)
        gather(plane, nn - 2 * plane)
        for side, n0 in ((0, 0), (1, nn - plane)):
            _lib.check(lib.fol_halo_add(s, h, side, step, n0, plane, d, _lib.ptr(R)))
        return ke_out, R
    if part.world == 1 or part.nz_local < 3:
        elements(0, ne)
        gather(0, nn)
        part.halo_sum(R, d, group)
        return ke_out, R
    # bottom-plane nodes touch only element layer 0, top-plane nodes only the last layer
    elements(0, layer)
    elements(ne - layer, layer)
    gather(0, plane)
    gather(nn - plane, plane)
    ready = torch.cuda.Event()
    ready.record()
    with torch.cuda.stream(comm_stream):
        comm_stream.wait_event(ready)
        part.halo_sum(R, d, group)
        done = torch.cuda.Event()
        done.record()
    # the interior element stage is a persistent kernel: leave a few thread blocks' worth of SM resources
    # free so the NCCL send/recv kernels of the side stream are scheduled next to it, not after it
    prev = lib.fol_set_grid_margin(GRID_MARGIN_CTAS)
    try:
        elements(layer, ne - 2 * layer)
    finally:
        lib.fol_set_grid_margin(prev)
    gather(plane, nn - 2 * plane)
    torch.cuda.current_stream().wait_event(done)
    return ke_out, R


def allreduce_gradients(parameters, group=None, bucket_bytes=64 << 20):
    """Sum `.grad` of the given parameters over all ranks with as few NCCL calls as fit the
    bucket size (flatten -> all_reduce -> unflatten)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0
    for g in grads:
        nbytes = g.numel() * g.element_size()
        if nbytes >= bucket_bytes // 4 and g.is_contiguous():
            # a large gradient (the output layer of a FOL network) is reduced in place: flattening it into a bucket
            # and copying it back would cost two extra passes over it
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            continue
        bucket.append(g)
        size += nbytes
        if size >= bucket_bytes:
            flush()
    flush()


class GradientReducer:
    """All-reduce of the network gradients OVERLAPPED with the backward pass: a post-accumulate hook on every
    parameter starts the NCCL all-reduce of its gradient on a side stream the moment autograd has produced it;
    `wait()` joins the side stream before the optimizer step.  Same sums as `allreduce_gradients` (one all-reduce per
    parameter, NCCL's fixed ring / tree order).

    A hook fires when autograd ACCUMULATES a gradient, i.e. after the whole backward node that produced it -- for a
    Linear layer after both its weight-gradient and its input-gradient GEMMs are enqueued.  For the (large) output
    layer of a FOL network that is too late: nothing sizeable is left of the backward pass to hide its 135 MB
    all-reduce behind.  `EarlyReduceLinear` is a Linear layer whose backward computes the weight / bias gradients
    FIRST, starts their all-reduce here (`reduce_now`) and only then computes the input gradient, which overlaps it;
    its parameters are passed in `exclude` so that no hook reduces them a second time.  On CPU tensors (gloo tests)
    everything runs synchronously."""

    def __init__(self, parameters, group=None, exclude=()):
        self.group = group
        self.active = dist.is_initialized() and dist.get_world_size(group) > 1
        self.stream = torch.cuda.Stream() if (self.active and torch.cuda.is_available()) else None
        self.handles = []
        skip = {id(p) for p in exclude}
        if self.active:
            for p in parameters:
                if p.requires_grad and id(p) not in skip:
                    self.handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def _start(self, tensors):
        if self.stream is None or not tensors[0].is_cuda:
            for t in tensors:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for t in tensors:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _hook(self, p):
        self._start([p.grad])

    def reduce_now(self, *tensors):
        """Start the in-place all-reduce of tensors that were just produced on the current stream (no-op when single
        rank); `wait()` must follow before they are read."""
        tensors = [t for t in tensors if t is not None]
        if self.active and tensors:
            self._start(tensors)

    def wait(self):
        if self.active and self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

    def close(self):
        for h in self.handles:
            h.remove()
        self.handles = []


class _EarlyReduceLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, reducer):
        ctx.save_for_backward(x, weight)
        ctx.reducer, ctx.has_bias = reducer, bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g2, x2 = g.reshape(-1, g.shape[-1]), x.reshape(-1, x.shape[-1])
        gw = g2.t().mm(x2) if ctx.needs_input_grad[1] else None          # (out, in): first, so that its all-reduce ...
        gb = g2.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        ctx.reducer.reduce_now(gw, gb)
        gx = g.matmul(w) if ctx.needs_input_grad[0] else None            # ... runs behind this GEMM
        ctx.reducer.wait()       # enqueued after the GEMM: gw / gb are complete before autograd accumulates them
        return gx, gw, gb, None


class EarlyReduceLinear(torch.nn.Linear):
    """torch.nn.Linear for the data-parallel FOL step: identical forward and gradients; in backward the weight / bias
    gradients are produced first and their all-reduce (through `reducer`, a GradientReducer that EXCLUDES this
    layer's parameters from its hooks) overlaps the input-gradient GEMM.  Without a reducer, or on one rank, it is a
    plain Linear."""

    def __init__(self, in_features, out_features, bias=True, reducer=None, **kw):
        super().__init__(in_features, out_features, bias=bias, **kw)
        self.reducer = reducer

    def forward(self, x):
        if self.reducer is None or not self.reducer.active:
            return super().forward(x)
        return _EarlyReduceLinearFn.apply(x, self.weight, self.bias, self.reducer)


def allreduce_loss_statistics(mean, stats, group=None):
    """Global (mean, (min, max, mean)) of `ComputeBatchLoss` when the samples are sharded evenly over the ranks
    (fe_loss.py:262 takes them over the whole batch): two tiny collectives on a packed 3-vector -- SUM for the mean,
    MAX for (-min, max).  Returns detached device scalars; the differentiable local mean stays the one to call
    `.backward()` on (its gradient, divided by the world size by `allreduce_gradients`' caller or summed as is, is
    the caller's choice of normalisation)."""
    mn, mx, _ = stats
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        m = mean.detach()
        return m, (mn.detach(), mx.detach(), m)
    world = dist.get_world_size(group)
    packed = torch.stack([mean.detach(), -mn.detach(), mx.detach()]).contiguous()
    total = packed[:1].clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    ext = packed[1:].clone()
    dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
    m = total[0] / world
    return m, (-ext[0], ext[1], m)


def shard_batch(n_samples, rank, world):
    """Contiguous, even split of the sample axis (the reference's P('data') sharding)."""
    if n_samples % world:
        raise ValueError("batch must be divisible by the number of ranks")
    per = n_samples // world
    return slice(rank * per, (rank + 1) * per)

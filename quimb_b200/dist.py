"""Multi-GPU sharding of the contraction path.

One process per GPU (torchrun); ``torch.distributed`` (NCCL over NVLink on
the GPUs, gloo in the CPU tests) is plumbing only.  What shards naturally on
this path are *independent units with a single exchange at the end*:

  * slices of a contraction tree: cotengra fixes the values of a few indices,
    every assignment is a complete independent contraction and the results
    are summed (quimb/tensor/tensor_core.py:255-259).  Units are dealt
    round-robin to the ranks, summed locally on the device and combined by
    ONE all-reduce of the (small) output -- no collective on the data path;
  * independent networks (many MPS norms / amplitudes): pure replicas.

The DMRG chain itself is sequential (each environment depends on the
previous one, dmrg.py:297-301), but its dominant cost -- the Lanczos matvecs
of the two-site eigensolve -- shards over the left bond with ONE exchange
step per matvec (``BondShard``): rank r owns the row slab a' in [lo_r, hi_r)
of the left environment L[a', w, a] and of every Krylov vector, all-gathers
the input vector (32 MiB at chi = 1024) and all-reduces the handful of inner
products of the re-orthogonalisation.  Everything else of the sweep (SVD,
environment updates) is replicated deterministically on every rank.
"""

import itertools

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_units(n_units, rank=None, world_size=None):
    """Round-robin assignment of ``n_units`` independent units."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(rank, n_units, world_size))


def all_reduce_sum(t):
    """In-place sum over ranks of a torch tensor (no-op without a group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def slice_assignments(sliced_inds, size_dict):
    """All value assignments of the sliced indices, in a fixed order."""
    ranges = [range(size_dict[ix]) for ix in sliced_inds]
    return list(itertools.product(*ranges))


def contract_sliced(arrays, inputs, output, sliced_inds=None, optimize="auto",
                    contract_fn=None, rank=None, world_size=None, reduce=True,
                    target_width=None, min_slices=None):
    """Slice-parallel contraction: sum over all assignments of ``sliced_inds``
    of the contraction with those indices fixed; assignments are sharded over
    the ranks and combined with one all-reduce.

    ``sliced_inds=None`` picks the indices with :func:`quimb_b200.tree.
    find_slices` (at least one slice per rank, intermediates of at most
    ``2**target_width`` elements).
    ``contract_fn(arrays, inputs, output, optimize)`` defaults to the device
    tree executor; the CPU tests inject a stand-in to exercise the sharding
    and the collective without a GPU.
    """
    if contract_fn is None:
        from .tree import array_contract as contract_fn
    inputs = [tuple(t) for t in inputs]
    if sliced_inds is None:
        # choose the slices from the tree: enough of them to feed every rank
        # and narrow enough to fit ``target_width`` (log2 elements)
        from .tree import find_slices, find_tree
        sd = {}
        for t, x in zip(inputs, arrays):
            for ix, d in zip(t, x.shape):
                sd[ix] = int(d)
        if min_slices is None:
            min_slices = world()[1] if world_size is None else world_size
        if optimize == "auto-hq" and target_width is not None:
            # tree and slices searched together (tree of the sliced network)
            from .tree import find_sliced_tree
            optimize, sliced_inds = find_sliced_tree(inputs, tuple(output), sd,
                                                     target_width, min_slices)
        else:
            full_tree = find_tree(inputs, tuple(output), sd, optimize)
            sliced_inds = find_slices(full_tree, target_width, min_slices)[0]
    if (rank is None and world_size is None and torch.distributed.is_available()
            and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1):
        # every rank searched for itself (deterministically: tree.py canonicalises
        # index names); rank 0's choice is nevertheless the one all ranks use
        from .tree import Tree
        steps = [(i, j) for i, j, _, _ in optimize.steps] if isinstance(optimize, Tree) else None
        box = [(tuple(sliced_inds), steps)]
        torch.distributed.broadcast_object_list(box, src=0)
        sliced_inds, steps0 = box[0]
        if steps0 is not None and isinstance(optimize, Tree) and steps0 != steps:
            red = [tuple(ix for ix in t if ix not in set(sliced_inds)) for t in inputs]
            optimize = Tree(red, tuple(output), optimize.size_dict, steps0)
    sliced = tuple(sliced_inds)
    if any(ix in output for ix in sliced):
        raise ValueError("cannot slice an output index")
    size_dict = {}
    for t, x in zip(inputs, arrays):
        for ix, d in zip(t, x.shape):
            size_dict[ix] = int(d)
    units = slice_assignments(sliced, size_dict)
    mine = shard_units(len(units), rank, world_size)
    red_inputs = [tuple(ix for ix in t if ix not in sliced) for t in inputs]
    if contract_fn.__module__.endswith("tree") and isinstance(optimize, (str, type(None))):
        # every slice has the same structure: find the tree once
        from .tree import find_tree
        optimize = find_tree(red_inputs, tuple(output), size_dict, optimize)
    total = None
    for u in mine:
        vals = dict(zip(sliced, units[u]))
        sub = []
        for t, x in zip(inputs, arrays):
            idx = tuple(vals[ix] if ix in vals else slice(None) for ix in t)
            sub.append(x[idx])
        part = contract_fn(sub, red_inputs, tuple(output), optimize)
        total = part if total is None else total + part
    if not reduce:
        return total, mine
    return _finish(total, arrays, inputs, output, size_dict), mine


def _finish(total, arrays, inputs, output, size_dict):
    from .array import Array
    t = total.t if isinstance(total, Array) else total
    if t is None:
        # this rank had no unit: contribute zeros of the output shape
        ref = arrays[0]
        rt = ref.t if isinstance(ref, Array) else ref
        shape = [size_dict[ix] for ix in output]
        t = torch.zeros(shape, dtype=rt.dtype if isinstance(rt, torch.Tensor) else torch.float64,
                        device=rt.device if isinstance(rt, torch.Tensor) else "cpu")
    elif not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    t = t.contiguous()
    all_reduce_sum(t)
    return t


class BondShard:
    """Row-slab sharding of a bond of size ``n`` over the ranks of a process
    group: the exchange layer of the sharded two-site eigensolve.

    NCCL moves device tensors directly over NVLink; with the gloo backend
    (CPU tests, or several ranks sharing one GPU in the single-GPU test) the
    collectives are staged through host memory.
    """

    def __init__(self, group=None, rank=None, world_size=None):
        self.group = group
        if rank is None or world_size is None:
            rank, world_size = world()
        self.rank, self.world_size = int(rank), int(world_size)
        self.active = (dist.is_available() and dist.is_initialized()
                       and self.world_size > 1)
        self._stage = self.active and dist.get_backend(group) != "nccl"
        self.bytes_gathered = 0

    def slab(self, n, rank=None):
        """[lo, hi) of this rank's rows of a bond of size ``n`` (balanced)."""
        r = self.rank if rank is None else rank
        return (n * r) // self.world_size, (n * (r + 1)) // self.world_size

    def all_reduce_(self, t):
        if not self.active:
            return t
        if self._stage and t.device.type != "cpu":
            h = t.cpu()
            dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.group)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_rows(self, local, n):
        """Concatenate the row slabs ``local`` (rows_r, cols) of all ranks into
        the full (n, cols) matrix, identical on every rank."""
        cols = local.shape[1]
        if not self.active:
            return local
        full = torch.empty((n, cols), dtype=local.dtype, device=local.device)
        self.bytes_gathered += full.numel() * full.element_size()
        even = n % self.world_size == 0
        if even and not self._stage:
            dist.all_gather_into_tensor(full, local.contiguous(), group=self.group)
            return full
        # ragged slabs / staged backend: pad every slab to the largest one
        rows = max(self.slab(n, r)[1] - self.slab(n, r)[0] for r in range(self.world_size))
        src = local.cpu() if self._stage else local
        pad = torch.zeros((rows, cols), dtype=src.dtype, device=src.device)
        pad[: src.shape[0]].copy_(src)
        parts = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(parts, pad, group=self.group)
        for r, part in enumerate(parts):
            lo, hi = self.slab(n, r)
            full[lo:hi].copy_(part[: hi - lo])
        return full


def mps_norm2_two_ended(sites, shape="lrp", group=None):
    """<psi|psi> with the chain contracted from both ends at once (SURVEY.md
    8e, cfg2): rank 0 carries the (chi, chi) environment in from the left over
    the first half of the sites, rank 1 from the right over the second half;
    ONE broadcast of the 8 MiB right environment (chi = 1024) and one
    (chi x chi) . (chi x chi) trace join them.  Each rank needs only its half
    of the site tensors resident (pass ``None`` for the others).  Beyond two
    ranks the chain does not shard this way (replicas only); without a process
    group both halves run locally.  Returns a 0-d device Array (identical on
    all ranks)."""
    from . import ops
    from .array import Array
    from .contract import contract_pair
    from .mps import norm_step, norm_step_right, site_lpr
    n = len(sites)
    if n < 2:
        from .mps import mps_norm2
        return mps_norm2(sites, shape)
    half = n // 2
    active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if active else 0

    def left():
        E = None
        for i in range(half):
            A = site_lpr(sites[i], shape, i, n)
            if E is None:
                E = ops.eye(A.shape[0], dtype=A.dtype, device=A.device)
            E = norm_step(E, A)
        return E

    def right():
        E = None
        for i in range(n - 1, half - 1, -1):
            A = site_lpr(sites[i], shape, i, n)
            if E is None:
                E = ops.eye(A.shape[2], dtype=A.dtype, device=A.device)
            E = norm_step_right(E, A)
        return E

    if not active:
        EL, ER = left(), right()
    else:
        EL = left() if rank == 0 else None
        ER = right() if rank == 1 else None
        stage = dist.get_backend(group) != "nccl"
        out = []
        for src, E in ((0, EL), (1, ER)):
            meta = [None]
            if rank == src:
                meta[0] = (tuple(E.shape), str(E.t.dtype))
            dist.broadcast_object_list(meta, src=src, group=group)
            shp, dt = meta[0]
            if rank == src:
                buf = ops.materialize(E, force=True).t
            else:
                buf = torch.empty(shp, dtype=getattr(torch, dt.split(".")[-1]),
                                  device=ops.default_device())
            wire = torch.view_as_real(buf) if buf.dtype.is_complex else buf
            if stage and wire.device.type != "cpu":
                host = wire.cpu()
                dist.broadcast(host, src=src, group=group)
                wire.copy_(host)
            else:
                dist.broadcast(wire, src=src, group=group)
            out.append(Array(buf))
        EL, ER = out
    # <psi|psi> = sum_{a', a} EL[a', a] ER[a', a]
    return Array(contract_pair(EL.t, [0, 1], ER.t, [0, 1], [], conj_a=EL.cj, conj_b=ER.cj))

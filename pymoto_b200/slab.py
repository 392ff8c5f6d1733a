"""z-slab domain decomposition: who owns which node planes / element layers, halo exchange, scalar all-reduce.

The reference is single-process; this is the multi-GPU layer of the new build (section 8e).
Node and element numbering are z-slowest (pymoto/common/domain.py:211,224), so a slab of node planes [k0, k1) is a
contiguous row range of K and of every nodal vector, and a contiguous range of element layers of the design.

One process per GPU; collectives go through ``torch.distributed`` (NCCL on GPUs; the same code runs on gloo + CPU
tensors, which is how the host-side logic is tested without a GPU).  The data path has exactly three exchange
steps: (1) neighbour halo planes before an operator application / transfer / filter, (2) a sum all-reduce of the
CG / LDAS dot products, (3) one gather of the first replicated multigrid level per V-cycle.
"""
import os

import torch
import torch.distributed as dist


class SlabPartition:
    """Ownership of node planes for every multigrid level.

    Level 0 is the finest grid with ``nz`` elements (nz+1 node planes) in z.  Rank r owns fine planes
    [r*m, (r+1)*m) with m = nz / world (the last rank also owns plane nz).  The first ``n_dist`` levels are split
    that way (level l: boundaries r*m / 2**l, which must be integers and even for the next coarser split level so
    that coarse plane K lives with fine plane 2K); all coarser levels are replicated on every rank.
    """

    def __init__(self, nz, world=1, rank=0, n_levels=1, min_planes=4, force_n_dist=None, level_dofs=None, min_dofs=0):
        self.nz, self.world, self.rank = int(nz), int(world), int(rank)
        if self.world > 1:
            if self.nz <= 0:
                raise ValueError("slab decomposition needs a 3-D grid")
            if self.nz % self.world != 0:
                raise ValueError(f"nz={nz} must be divisible by the number of ranks {world}")
        self.m = self.nz // self.world if self.world > 1 else self.nz
        # number of slab-distributed levels: planes per rank stay >= min_planes and boundaries stay integral
        # (level l may be split only if its boundaries r*m/2**l are even, so that coarse plane K lives with fine plane 2K)
        n_dist = 1
        if self.world > 1:
            if n_levels > 1 and self.m % 2 != 0:
                raise ValueError(f"nz / ranks = {self.m} must be even for a multigrid hierarchy")
            # the coarsest level (index n_levels-1) is solved directly and is always replicated
            while (n_dist < n_levels - 1 and self.m % (2 ** (n_dist + 1)) == 0 and self.m // (2 ** n_dist) >= min_planes
                   and (level_dofs is None or level_dofs[n_dist] >= min_dofs)):
                n_dist += 1
        else:
            n_dist = n_levels
        if force_n_dist is not None:
            n_dist = force_n_dist
        self.n_dist = n_dist
        self.n_levels = n_levels

    def is_distributed(self, level):
        return self.world > 1 and level < self.n_dist

    def planes(self, level, rank=None):
        """Owned node planes [k0, k1) of ``rank`` at ``level`` (whole grid for replicated levels)."""
        r = self.rank if rank is None else rank
        nzl = self.nz >> level
        if not self.is_distributed(level):
            return 0, nzl + 1
        ml = self.m >> level
        k0 = r * ml
        k1 = (r + 1) * ml if r < self.world - 1 else nzl + 1
        return k0, k1

    def slab_planes(self, level, rank=None):
        """Planes [k0, k1) rank would own at ``level`` if that level were split (used at the transition to the first
        replicated level: each rank still produces the coarse rows that sit on its fine slab)."""
        r = self.rank if rank is None else rank
        nzl = self.nz >> level
        if self.world == 1:
            return 0, nzl + 1
        ml = self.m >> level
        return r * ml, ((r + 1) * ml if r < self.world - 1 else nzl + 1)

    def elem_layers(self, level=0, rank=None):
        """Owned element layers [e0, e1): layer e belongs to the owner of node plane e."""
        k0, k1 = self.planes(level, rank)
        return k0, min(k1, self.nz >> level)

    @property
    def lower(self):
        return self.rank - 1 if self.rank > 0 else None

    @property
    def upper(self):
        return self.rank + 1 if self.rank < self.world - 1 else None


class SlabComm:
    """Neighbour halo exchange and small collectives over a torch.distributed process group."""

    def __init__(self, part: SlabPartition, group=None):
        self.part = part
        self.group = group
        self.exchanges = 0
        self.allreduces = 0
        self.fast = False  # symmetric-memory mailboxes enabled (CUDA + NCCL only)
        self.fused = False  # one-launch exchange kernels (pmb_peer_*) on top of them
        self.fast_exchanges = 0
        self.fast_allreduces = 0
        self._fx_red_max = 16 if os.environ.get("PMB_PEER_ALLREDUCE", "1") != "0" else 0

    def enable_mailboxes(self, max_doubles, device):
        """Halo exchange through symmetric memory instead of NCCL send/recv: every rank gets a mailbox of
        2 slots x 2 directions x ``max_doubles``; neighbours store their boundary planes into it with a copy kernel
        (peer stores over NVLink), a device-side barrier orders the stores before the unpack.  Two slots alternate so
        that one barrier per exchange is enough (a slot is rewritten two exchanges later, i.e. after another barrier
        that the receiver only reaches once it has unpacked)."""
        if not self.active:
            return False
        import torch.distributed._symmetric_memory as symm_mem

        self._cap = int(max_doubles)
        # 4 slots x 2 directions: slots 0 / 1 alternate for eagerly launched exchanges, slots 2 / 3 for exchanges captured in a
        # CUDA graph (a replayed graph always starts on slot 2, so it can never reuse the slot of the exchange just before it)
        self._mb = symm_mem.empty(8 * self._cap, dtype=torch.float64, device=device)
        self._mb.zero_()
        self._symm = symm_mem
        self._hdl = symm_mem.rendezvous(self._mb, dist.group.WORLD if self.group is None else self.group)
        p = self.part
        self._peer_lo = self._hdl.get_buffer(p.lower, (8 * self._cap,), torch.float64) if p.lower is not None else None
        self._peer_hi = self._hdl.get_buffer(p.upper, (8 * self._cap,), torch.float64) if p.upper is not None else None
        self._slot = 0
        self._gslot = 0
        self.fast = True
        self.fused = False
        if os.environ.get("PMB_PEER_FUSED", "1") != "0" and p.world <= 16:
            self._enable_fused(device)
        return True

    def _enable_fused(self, device):
        """One-launch exchange steps (``pmb_peer_halo_exchange`` / ``pmb_peer_allreduce``): a second symmetric allocation holds
        2 mailbox slots x 2 directions and the tables of the small all-reduce (every double travels as a 16-byte unit that
        carries the exchange number, so there are no separate flags and no barrier); the kernels number the exchanges themselves
        (a counter in local device memory), so eager launches and CUDA-graph replays share the slots."""
        import ctypes as C

        from . import _lib

        p, cap, W = self.part, self._cap, self.part.world
        n_box, n_tab = _lib.query("pmb_peer_halo_box_doubles", cap), _lib.query("pmb_peer_reduce_table_doubles", W)
        total = n_box + n_tab
        self._fx = self._symm.empty(total, dtype=torch.float64, device=device)  # all-zero bits = exchange number 0 in every unit
        self._fx.zero_()
        self._fx_ctl = torch.zeros(8, dtype=torch.int64, device=device)
        torch.cuda.synchronize()
        self._fx_hdl = self._symm.rendezvous(self._fx, dist.group.WORLD if self.group is None else self.group)
        dist.barrier(group=self.group)  # nobody stores into a peer's mailbox before that peer has zeroed it
        base = [self._fx_hdl.get_buffer(r, (total,), torch.float64).data_ptr() for r in range(W)]
        self._fx_halo = _lib.PeerHalo(base[p.rank], base[p.lower] if p.lower is not None else None,
                                      base[p.upper] if p.upper is not None else None, self._fx_ctl.data_ptr(), cap)
        red = _lib.PeerReduce()
        red.world, red.rank, red.ctl = W, p.rank, self._fx_ctl.data_ptr() + 8 * 4
        for r in range(W):
            red.slots[r] = base[r] + 8 * n_box
        self._fx_red = red
        self._fx_byref = (C.byref(self._fx_halo), C.byref(self._fx_red))
        self.fused = True

    def check_peer_timeouts(self):
        """Synchronise and raise if a one-launch exchange kernel ever gave up waiting for a peer (its results are invalid)."""
        if self.fused:
            torch.cuda.synchronize()
            ctl = self._fx_ctl.cpu()
            if int(ctl[2]) or int(ctl[5]):
                from . import _lib

                raise _lib.PmbError(f"rank {self.part.rank}: a peer never arrived (halo exchange #{int(ctl[2])}, all-reduce "
                                    f"#{int(ctl[5])} timed out): the ranks issued different exchange sequences, or a rank died")

    # ---- CUDA-graph capture of code that exchanges halos (the slab V-cycle)
    def begin_capture(self):
        self._gslot = 0

    def end_capture(self):
        """Last captured operation: one more barrier, so that two back-to-back replays (last exchange of one, first of the
        next) are separated like any two exchanges on different slots are."""
        if self.fast:
            self._hdl.barrier(channel=0)

    def symmetric_vector(self, n):
        """A zeroed float64 vector in symmetric memory and the views of every rank's copy (for :meth:`bcast_part`)."""
        buf = self._symm.empty(int(n), dtype=torch.float64, device=self._mb.device)
        buf.zero_()
        hdl = self._symm.rendezvous(buf, dist.group.WORLD if self.group is None else self.group)
        peers = [hdl.get_buffer(r, (int(n),), torch.float64) for r in range(self.part.world)]
        return buf, peers, hdl

    def bcast_part(self, local, peers, hdl, offset):
        """Replicate a distributed vector without a collective: every rank stores its part at ``offset`` of every rank's
        copy (peer stores over NVLink), one device-side barrier completes the vector everywhere.  Deterministic, capturable."""
        from . import _lib

        st = torch.cuda.current_stream().cuda_stream
        n, src = local.numel(), local.data_ptr()
        dsts = [p.data_ptr() + 8 * offset for p in peers]
        if self.fused:  # the one-launch halo exchanges synchronise neighbours only: before overwriting every rank's copy make
            hdl.barrier(channel=0)  # sure every rank is done reading the previous one (the old exchanges were global barriers)
        for i in range(0, len(dsts), 2):
            _lib.call("pmb_halo_copy2", n, src, dsts[i], src if i + 1 < len(dsts) else None, dsts[i + 1] if i + 1 < len(dsts) else None, st)
        hdl.barrier(channel=0)
        self.allreduces += 1

    def _fast_exchange(self, base, own_offset, own_len, n, lower, upper):
        from . import _lib

        p = self.part
        st = torch.cuda.current_stream().cuda_stream
        b0 = base.data_ptr() + 8 * own_offset  # first owned entry
        if self.fused:  # pack, publish, wait, unpack in one launch; `upper` = my upper halo is wanted = bottom planes travel down
            _lib.call("pmb_peer_halo_exchange", self._fx_byref[0], n, b0 if upper else None, b0 + 8 * (own_len - n) if lower else None,
                      b0 - 8 * n if lower else None, b0 + 8 * own_len if upper else None, st)
            self.exchanges += 1
            self.fast_exchanges += 1
            return
        if torch.cuda.is_current_stream_capturing():
            o = (2 + (self._gslot & 1)) * 2 * self._cap
            self._gslot += 1
        else:
            o = (self._slot & 1) * 2 * self._cap
            self._slot += 1
        # box layout per slot: [0, cap) = planes coming from the rank below, [cap, 2 cap) = from the rank above
        # `upper` = I want my upper halo filled = every rank sends its bottom planes down; `lower` = top planes go up
        src0 = dst0 = src1 = dst1 = None
        if upper and self._peer_lo is not None:
            src0, dst0 = b0, self._peer_lo.data_ptr() + 8 * (o + self._cap)
        if lower and self._peer_hi is not None:
            src1, dst1 = b0 + 8 * (own_len - n), self._peer_hi.data_ptr() + 8 * o
        _lib.call("pmb_halo_copy2", n, src0, dst0, src1, dst1, st)
        self._hdl.barrier(channel=0)
        src0 = dst0 = src1 = dst1 = None
        if lower and p.lower is not None:
            src0, dst0 = self._mb.data_ptr() + 8 * o, b0 - 8 * n
        if upper and p.upper is not None:
            src1, dst1 = self._mb.data_ptr() + 8 * (o + self._cap), b0 + 8 * own_len
        _lib.call("pmb_halo_copy2", n, src0, dst0, src1, dst1, st)
        self.exchanges += 1
        self.fast_exchanges += 1

    @property
    def active(self):
        return self.part.world > 1

    def exchange(self, base, own_offset, own_len, plane, lower=True, upper=True, width=1):
        """Fill the halo planes of a padded 1-D buffer ``base`` (any dtype).

        ``base[own_offset : own_offset + own_len]`` are the owned entries, ``plane`` entries per plane; the ``width``
        planes below / above them are halos.  ``lower``: receive my lower halo (the lower neighbour sends its top
        owned planes); ``upper``: receive my upper halo.  Every rank must call this with the same flags.
        """
        if not self.active:
            return
        p = self.part
        n = plane * width
        if self.fast and base.is_cuda and base.dtype == torch.float64 and n <= self._cap:
            return self._fast_exchange(base, own_offset, own_len, n, lower, upper)
        ops = []
        if upper:  # data flows downwards: I send my bottom planes to the lower neighbour, receive from the upper
            if p.lower is not None:
                ops.append(dist.P2POp(dist.isend, base[own_offset:own_offset + n], p.lower, self.group))
            if p.upper is not None:
                ops.append(dist.P2POp(dist.irecv, base[own_offset + own_len:own_offset + own_len + n], p.upper, self.group))
        if lower:  # data flows upwards
            if p.upper is not None:
                ops.append(dist.P2POp(dist.isend, base[own_offset + own_len - n:own_offset + own_len], p.upper, self.group))
            if p.lower is not None:
                ops.append(dist.P2POp(dist.irecv, base[own_offset - n:own_offset], p.lower, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.exchanges += 1

    def exchange_start(self, base, own_offset, own_len, plane):
        """Post the halo exchange (both sides, one plane) without waiting: returns the request handles.  The transfers
        run on the communicator's own stream, ordered after everything already queued on the current stream."""
        if not self.active:
            return []
        p = self.part
        ops = []
        if p.lower is not None:
            ops.append(dist.P2POp(dist.isend, base[own_offset:own_offset + plane], p.lower, self.group))
            ops.append(dist.P2POp(dist.irecv, base[own_offset - plane:own_offset], p.lower, self.group))
        if p.upper is not None:
            ops.append(dist.P2POp(dist.isend, base[own_offset + own_len - plane:own_offset + own_len], p.upper, self.group))
            ops.append(dist.P2POp(dist.irecv, base[own_offset + own_len:own_offset + own_len + plane], p.upper, self.group))
        self.exchanges += 1
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def exchange_finish(reqs):
        """Make the current stream wait for a posted exchange."""
        for r in reqs:
            r.wait()

    def allreduce_(self, t, op="sum"):
        """In-place sum (or maximum) over the slabs.  Up to 16 doubles on the device go through the one-launch peer-memory
        kernel (combined in rank order: identical bits on every rank), everything else through the process group."""
        if self.active:
            if (getattr(self, "fused", False) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
                    and 1 <= t.numel() <= self._fx_red_max):
                from . import _lib

                _lib.call("pmb_peer_allreduce", self._fx_byref[1], t.data_ptr(), t.numel(), 0 if op == "sum" else 1,
                          torch.cuda.current_stream().cuda_stream)
                self.fast_allreduces += 1
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=self.group)
            self.allreduces += 1
        return t

    def gather_full(self, local, full, offset):
        """Replicate a distributed array: every rank writes its owned part at ``offset`` of the zeroed ``full``
        buffer, then a sum all-reduce (sizes at the first replicated level are small)."""
        full.zero_()
        full[offset:offset + local.numel()] = local
        return self.allreduce_(full)

    def barrier(self):
        if self.active:
            dist.barrier(group=self.group)


def world_from_env():
    """(rank, world) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class SlabContext:
    """Process-wide decomposition state: the partition of the finest grid and the communicator.

    Set once per process with :func:`init` (like torch.distributed's default group) so that the module signatures
    stay exactly the reference's; without it every module runs single-GPU on the whole grid.
    """

    def __init__(self, part: SlabPartition, comm: SlabComm):
        self.part, self.comm = part, comm

    @property
    def active(self):
        return self.part.world > 1


_context = None


def init(domain, n_levels=1, group=None, min_planes=4, force_n_dist=None, ndof=3, min_dofs=None, mailboxes=True):
    """Decompose ``domain`` in z over the ranks of the (default) process group. Call after init_process_group.

    ``n_levels`` = number of matrices in the multigrid hierarchy (GeometricMultigrid operators + 1).  Levels with
    fewer than ``min_dofs`` unknowns (default 200 000, ``PMB_SLAB_MIN_DOFS``: with 4-7 us per one-launch halo exchange a level
    of that size is cheaper split than replicated, measured) or fewer than ``min_planes`` node planes per rank are replicated
    on every rank.
    """
    global _context
    if min_dofs is None:
        min_dofs = int(os.environ.get("PMB_SLAB_MIN_DOFS", 200_000))
    rank, world = world_from_env()
    nx, ny, nz = int(domain.nelx), int(domain.nely), int(getattr(domain, "nelz", 0) or 0)
    level_dofs = [ndof * ((nx >> l) + 1) * ((ny >> l) + 1) * ((nz >> l) + 1) for l in range(max(n_levels, 1))]
    part = SlabPartition(nz, world, rank, n_levels=n_levels, min_planes=min_planes, force_n_dist=force_n_dist,
                         level_dofs=level_dofs, min_dofs=min_dofs)
    comm = SlabComm(part, group)
    _context = SlabContext(part, comm)
    if (mailboxes and world > 1 and torch.cuda.is_available() and dist.get_backend(group) == "nccl"
            and os.environ.get("PMB_HALO_MAILBOX", "1") != "0"):
        try:  # nodal-vector planes of every level fit the finest level's plane
            comm.enable_mailboxes((nx + 1) * (ny + 1) * ndof, torch.device("cuda", torch.cuda.current_device()))
        except Exception as e:  # symmetric memory unavailable (no P2P access, old driver): NCCL send/recv stays in use
            import warnings

            warnings.warn(f"symmetric-memory halo mailboxes unavailable ({e}); using NCCL send/recv")
    return _context


def reset():
    global _context
    _context = None


def context(nz=None, n_levels=1):
    """The active context, or a trivial single-rank one.  ``nz`` (when given) must be the z-size the active decomposition
    was made for: a module built on another domain would silently use the wrong slab."""
    if _context is not None:
        if nz is not None and _context.active and int(nz) != _context.part.nz:
            raise ValueError(f"active slab decomposition is for nz = {_context.part.nz}, module domain has nz = {nz}; "
                             "call pymoto_b200.slab.init(domain, ...) for this domain (or slab.reset())")
        return _context
    part = SlabPartition(nz or 0, 1, 0, n_levels=n_levels)
    return SlabContext(part, SlabComm(part))

"""Slab decomposition of the per-timestep advance across the GPUs of one node (one process per GPU).

The domain is split along x (i): rows are contiguous `ydim` runs in the reference layout, so a slab's halo is
2 contiguous rows per evolved plane and side.  Per RK stage each rank
    1. runs the fused stage kernel on its slab (spruce_mgpu_stage), which also packs its first/last 2 rows of the
       8 evolved planes into two contiguous staging buffers,
    2. exchanges the staging buffers with its ring neighbours (torch.distributed isend/irecv -> NCCL send/recv over
       NVLink; the buffers are library memory aliased as torch tensors, no copies),
    3. unpacks the received rows into its halo rows (spruce_mgpu_unpack);
after the last stage the dt minimum (one double) is all-reduced with MIN.  min/max are exact, so N-GPU runs
reproduce the 1-GPU run bit for bit (tests/test_gpu_parity.py::test_two_slabs...).

Two transports (SlabRunner(transport=...)):
  "p2p"  (default) the library's own peer-store exchange: the pack kernel writes the edge rows into the neighbours' memory over
         NVLink (CUDA IPC mapping) and signals with a sequence number; torch.distributed only carries the 64-byte IPC handles at
         start-up.  The whole time loop is one spruce_advance call per rank.
  "nccl" the caller-owned exchange described above (batch_isend_irecv + all_reduce), kept as the reference transport.

The partition / neighbour / exchange logic is backend independent (`exchange_halos`) and is covered on CPU with
gloo and world_size 2 (tests/test_multigpu_host.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

HALO = 2


def partition(xdim: int, world: int):
    """Row ranges [row0, row0+n) per rank: as even as possible, every slab at least 2*HALO rows."""
    base, rem = divmod(xdim, world)
    out, r0 = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((r0, n))
        r0 += n
    if any(n < 2 * HALO for _, n in out):
        raise ValueError("xdim=%d is too small for %d slabs (need >= %d rows each)" % (xdim, world, 2 * HALO))
    return out


def neighbours(rank: int, world: int, periodic_x: bool):
    """(lower neighbour, upper neighbour) or None at a physical (non-periodic) boundary."""
    lo = rank - 1 if rank > 0 else (world - 1 if periodic_x else None)
    hi = rank + 1 if rank < world - 1 else (0 if periodic_x else None)
    return lo, hi


def exchange_halos(dist, send_lo, send_hi, recv_lo, recv_hi, rank: int, world: int, periodic_x: bool):
    """Ring exchange of the packed halo buffers: my first rows go to the lower neighbour's upper halo, my last rows to
    the upper neighbour's lower halo.  All four transfers are posted as one batch (one NCCL group)."""
    lo, hi = neighbours(rank, world, periodic_x)
    ops = []
    if lo is not None:
        ops.append(dist.P2POp(dist.isend, send_lo, lo, tag=1))
        ops.append(dist.P2POp(dist.irecv, recv_lo, lo, tag=2))
    if hi is not None:
        ops.append(dist.P2POp(dist.isend, send_hi, hi, tag=2))
        ops.append(dist.P2POp(dist.irecv, recv_hi, hi, tag=1))
    if world == 2 and periodic_x and lo == hi:
        # both neighbours are the same rank: order the pairs identically on both sides (lo-send matches the peer's hi-recv)
        ops = [dist.P2POp(dist.isend, send_lo, lo, tag=1), dist.P2POp(dist.irecv, recv_hi, hi, tag=1),
               dist.P2POp(dist.isend, send_hi, hi, tag=2), dist.P2POp(dist.irecv, recv_lo, lo, tag=2)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class _DevBuf:
    """Library-owned device memory exposed through __cuda_array_interface__ so torch can alias it (no copy)."""

    def __init__(self, ptr: int, n: int, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None}


class SlabRunner:
    def __init__(self, planes, ion_mass, adiabatic_index, *, rank, world, device, transport="p2p", xdim=None, **kw):
        import torch
        import torch.distributed as dist
        from . import capi
        from .domain import PlasmaDomain
        self.torch, self.dist, self.capi = torch, dist, capi
        self.rank, self.world = rank, world
        # `planes` holds either global planes or (xdim given) only this rank's rows; d_x is always the global 1-D array then
        if xdim is None:
            if planes["d_x"].ndim != 2:
                raise ValueError("xdim is required when the planes hold only this rank's rows")
            xdim = planes["d_x"].shape[0]
        self.transport = transport
        self.row0, self.nx = partition(xdim, world)[rank]
        self.periodic_x = kw.get("xb", ("periodic", "periodic")) == ("periodic", "periodic")
        self.dom = PlasmaDomain(planes, ion_mass, adiabatic_index, device=device, row0=self.row0, nx_local=self.nx,
                                rank=rank, n_ranks=world, setup=False, **kw)
        self.lib = self.dom.lib
        self.stream = torch.cuda.ExternalStream(self.dom.stream())
        self._nccl_ready = False
        # setup = local populateVariablesFromState, then halos of the primary state and the global dt minimum
        # zero-plane knowledge must be global (a neighbour's halo rows may be non-zero): OR the per-rank masks
        lm = C.c_int()
        capi.check(self.lib.spruce_plane_activity(self.dom.h, C.byref(lm), -1))
        bits = torch.tensor([(lm.value >> b) & 1 for b in range(5)], dtype=torch.int32, device="cuda")
        dist.all_reduce(bits, op=dist.ReduceOp.MAX)
        gm = sum(int(v) << b for b, v in enumerate(bits.tolist()))
        capi.check(self.lib.spruce_plane_activity(self.dom.h, None, gm))
        if transport == "p2p":
            mine = (C.c_char * 64)()
            ok = self.lib.spruce_mgpu_ipc_export(self.dom.h, mine) == 0
            t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).cuda()
            allh = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allh, t)
            blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh)
            ok = ok and self.lib.spruce_mgpu_ipc_connect(self.dom.h, blob, world) == 0
            # every rank must have mapped its peers (CUDA IPC needs peer access between the devices and a shared IPC namespace);
            # otherwise ALL ranks switch to the NCCL transport together
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if rank == 0:
                    import sys
                    print("spruce_b200: CUDA IPC peer mapping unavailable (%s); using the NCCL halo transport" % capi.load().spruce_last_error().decode(), file=sys.stderr)
                transport = self.transport = "nccl"
        if transport == "p2p":
            self.dom.setup()
            dist.barrier()                      # every rank has mapped its neighbours and finished its local setup
            capi.check(self.lib.spruce_mgpu_initial_exchange(self.dom.h))
            return
        self._init_nccl_buffers()
        self.dom.setup()
        with torch.cuda.stream(self.stream):
            capi.check(self.lib.spruce_mgpu_pack(self.dom.h, 3))        # static planes: be_* are transported and need halo rows
            self._exchange()
            capi.check(self.lib.spruce_mgpu_unpack(self.dom.h, 3))
            capi.check(self.lib.spruce_mgpu_pack(self.dom.h, 0))
            self._exchange()
            capi.check(self.lib.spruce_mgpu_unpack(self.dom.h, 0))
            dist.all_reduce(self.dtmin, op=dist.ReduceOp.MIN)
        self.dom.synchronize()

    def _init_nccl_buffers(self):
        """Tensors aliasing the library's packed halo buffers and dt word, for the caller-driven (NCCL) transport."""
        torch, capi = self.torch, self.capi
        ptrs = [C.c_void_p() for _ in range(4)]
        nbytes = C.c_size_t()
        capi.check(self.lib.spruce_halo_buffers(self.dom.h, *[C.byref(p) for p in ptrs], C.byref(nbytes)))
        n = nbytes.value // 8
        self.send_lo, self.send_hi, self.recv_lo, self.recv_hi = [torch.as_tensor(_DevBuf(p.value, n), device="cuda") for p in ptrs]
        p = C.c_void_p()
        capi.check(self.lib.spruce_mgpu_dt_min_ptr(self.dom.h, C.byref(p)))
        self.dtmin = torch.as_tensor(_DevBuf(p.value, 1), device="cuda")
        ns = C.c_int()
        capi.check(self.lib.spruce_mgpu_n_stages(self.dom.h, C.byref(ns)))
        self.n_stages = ns.value
        self.out_set = []
        for s in range(self.n_stages):
            w = C.c_int()
            capi.check(self.lib.spruce_mgpu_stage_output(self.dom.h, s, C.byref(w)))
            self.out_set.append(w.value)
        self._nccl_ready = True

    def _exchange(self):
        exchange_halos(self.dist, self.send_lo, self.send_hi, self.recv_lo, self.recv_hi, self.rank, self.world, self.periodic_x)

    def step(self, n_steps: int = 1):
        """n_steps x advanceTime on the slab; everything is stream ordered (no host synchronisation inside)."""
        capi, lib, h, dist = self.capi, self.lib, self.dom.h, self.dist
        if self.transport == "p2p":
            return self.dom.advance(n_steps)
        with self.torch.cuda.stream(self.stream):
            for _ in range(n_steps):
                capi.check(lib.spruce_mgpu_begin_step(h))
                for s in range(self.n_stages):
                    capi.check(lib.spruce_mgpu_stage(h, s))
                    self._exchange()
                    capi.check(lib.spruce_mgpu_unpack(h, self.out_set[s]))
                dist.all_reduce(self.dtmin, op=dist.ReduceOp.MIN)
                capi.check(lib.spruce_mgpu_end_step(h))

    def gather(self, name: str):
        """Rank 0 gets the global plane (rows concatenated in rank order); other ranks get None."""
        torch, dist = self.torch, self.dist
        local = torch.from_numpy(self.dom.grid(name)).cuda()
        parts = partition(self.dom.xdim, self.world)
        if self.rank == 0:
            bufs = [torch.empty((n, self.dom.ydim), dtype=torch.float64, device="cuda") for _, n in parts]
            bufs[0].copy_(local)
            for r in range(1, self.world):
                dist.recv(bufs[r], src=r)
            return torch.cat(bufs).cpu().numpy()
        dist.send(local, dst=0)
        return None

    def bench(self, steps: int, warmup: int):
        """Contract timing: barrier + synchronize on both sides, CUDA events on the stream, MAX over ranks."""
        torch, dist = self.torch, self.dist
        self.step(warmup)
        self.dom.synchronize()
        l0 = self.dom.launch_count()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        self.step(steps)
        e1.record(self.stream)
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launches = torch.tensor([self.dom.launch_count() - l0], dtype=torch.int64, device="cuda")
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        cells = self.dom.xdim * self.dom.ydim
        t = self.dom.time
        return dict(value=cells * steps / (ms.item() * 1e-3), ms=ms.item(), launches=int(launches.item()), time=t,
                    clocks=None, roofline=None, e2e=None, cpu=None)

    def close(self):
        self.dom.close()

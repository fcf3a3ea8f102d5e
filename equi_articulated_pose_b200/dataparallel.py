"""Data-parallel plumbing: one process per GPU, clouds sharded over ranks, gradient all-reduce.

The reference trains with DistributedDataParallel over NCCL, one cloud per rank
(SPConvNets/trainer_unsup_arti_align.py:52-56,203-208,430-440): the convolution is per-sample,
so the only exchange step of the path is the gradient all-reduce (7.67 M fp32 = 30.7 MB for the
classic backbone).  Here the gradients are packed into ONE flat fp32 bucket after backward (one
concatenation kernel; the parameters' .grad tensors become views of it), so a step needs exactly one
NCCL all-reduce (NVLink 5 / NVSwitch, NVLS in-switch reduction when available) and the optimiser
works on the flat storage.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (torchrun).  -> (rank, local_rank, world)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(total, rank, world):
    """Contiguous [lo, hi) slice of `total` clouds owned by `rank` (DistributedSampler-like, no padding)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradBucket:
    """All gradients of `params` in one contiguous fp32 buffer, gathered AFTER backward.

    `zero_()` drops the gradients (`p.grad = None`), so autograd hands every parameter its freshly computed gradient tensor
    instead of launching one `grad += new` kernel per parameter (59 small launches per step on the classic backbone);
    `collect()` then packs them into the flat buffer with ONE concatenation kernel and re-points every `p.grad` at its
    slice, so the optimiser and the all-reduce work on the flat storage.  `flat` collects on first access."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self._flat = torch.zeros(total, dtype=torch.float32, device=ref.device)
        self._views, off = [], 0
        for p in self.params:
            n = p.numel()
            self._views.append(self._flat[off:off + n].view_as(p))
            off += n
        for p, v in zip(self.params, self._views):
            p.grad = v
        self._collected = True

    @property
    def flat(self):
        self.collect()
        return self._flat

    def zero_(self):
        """Forget the gradients of the previous step (the next backward writes new tensors; nothing is memset)."""
        for p in self.params:
            p.grad = None
        self._collected = False

    def collect(self):
        """Pack the parameters' gradients into the flat buffer (one kernel) and make every `.grad` a view of it."""
        if self._collected and all(p.grad is v or (p.grad is not None and p.grad.data_ptr() == v.data_ptr())
                                   for p, v in zip(self.params, self._views)):
            return
        base = self._flat.untyped_storage().data_ptr()
        pieces, aliased = [], False
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is None:
                g = torch.zeros_like(v)                          # parameter not reached by this backward
            elif g.untyped_storage().data_ptr() == base:
                aliased = True                                   # still one of our views (accumulated in place)
            pieces.append(g.reshape(-1))
        if aliased:                                              # mixed state: never concatenate a buffer into itself
            pieces = [t.clone() for t in pieces]
        torch.cat(pieces, out=self._flat)
        for p, v in zip(self.params, self._views):
            p.grad = v
        self._collected = True

    def all_reduce_mean(self, group=None):
        """One collective for the whole model; averages over ranks like DDP does."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return                              # one rank: nothing to exchange, the gradients stay where backward wrote them
        self.collect()
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=group)
        self._flat.div_(dist.get_world_size(group))

    def nbytes(self):
        return self._flat.numel() * 4


class PeerMailbox:
    """SyncBatchNorm exchanges over NVLink peer memory (csrc/peer.cu) instead of NCCL all-reduces.

    Every rank cudaMallocs a mailbox, the IPC handles are exchanged once over the process group, every rank maps its
    peers' mailboxes; after that an exchange is one single-CTA kernel (stores into the peers' HBM + flags).
    Construction is collective and returns a usable object on ALL ranks or raises on all of them (the outcome is
    agreed with one all-reduce), so callers can fall back to NCCL consistently."""

    def __init__(self, device, group=None):
        import ctypes
        from . import lib
        self.device, self.group = device, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.seq = 0
        self._own, self._opened = None, []
        ok, err = 1, ""
        handle = bytes(64)
        try:
            nbytes = ctypes.c_int64(0)
            lib.setup_call("vgtkb_peer_mailbox_bytes", device, self.world, ctypes.byref(nbytes))
            own = ctypes.c_void_p(0)
            hbuf = ctypes.create_string_buffer(64)
            lib.setup_call("vgtkb_peer_alloc", device, nbytes.value, ctypes.byref(own), ctypes.cast(hbuf, ctypes.c_void_p))
            self._own, handle = own.value, hbuf.raw
        except Exception as e:                      # agreed below
            ok, err = 0, str(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, handle), group=group)
        ptrs = [None] * self.world
        if all(h[0] for h in handles):
            try:
                for r, (_, h) in enumerate(handles):
                    if r == self.rank:
                        ptrs[r] = self._own
                    else:
                        q = ctypes.c_void_p(0)
                        hb = ctypes.create_string_buffer(h, 64)
                        lib.setup_call("vgtkb_peer_open", device, ctypes.cast(hb, ctypes.c_void_p), ctypes.byref(q))
                        ptrs[r] = q.value
                        self._opened.append(q.value)
            except Exception as e:
                ok, err = 0, str(e)
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close()
            raise RuntimeError(f"peer mailboxes unavailable on at least one rank ({err or 'another rank failed'})")
        self.ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        dist.barrier(group=group)

    def next_seq(self):
        """Sequence argument of the exchange kernels: 0 = the kernel takes the next value of the mailbox's device-side
        counter, so the call can be captured in a CUDA graph and replayed (every rank runs the same exchanges)."""
        return 0

    def all_reduce_sums(self, buf):
        """buf (fp64, <= 2049 elements) <- sum over ranks, in place, on the current stream."""
        from . import lib
        lib.call("vgtkb_peer_allreduce_f64", self.device, buf.numel(), lib.ptr(buf), self.rank, self.world, self.ptrs,
                 self.next_seq())

    def status(self):
        """0, or r + 1 when rank r missed an exchange (timeout VGTKB_PEER_TIMEOUT_S, default 600 s).  Synchronises."""
        import ctypes
        from . import lib
        st = ctypes.c_int64(0)
        lib.call("vgtkb_peer_status", self.device, self._own, self.world, ctypes.byref(st))
        return int(st.value)

    def close(self):
        from . import lib
        for q in self._opened:
            try:
                lib.setup_call("vgtkb_peer_close", self.device, q)
            except Exception:
                pass
        self._opened = []
        if self._own:
            try:
                lib.setup_call("vgtkb_peer_free", self.device, self._own)
            except Exception:
                pass
            self._own = None

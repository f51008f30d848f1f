"""Multi-process plumbing for jobs that run one process per GPU (bench.py, launched by torchrun).

The data path never communicates: every rank owns a contiguous slab of independent channels
(all hot-path effects are channel-wise, EFFECT_FLAG_CH_DEPS_IDENTITY) and its own copy of the
operator state.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only for the barrier
around the timed region and for the max-/sum-reduction of per-rank scalars.
"""
import os


def channel_slab(total_channels, parts, index):
    """Contiguous channel range [begin, end) of slab `index` out of `parts` -- the same rule
    dspb200_chain_create() uses for its shards (api.cu: ch_begin = channels * i / n_shards)."""
    begin = total_channels * index // parts
    end = total_channels * (index + 1) // parts
    return begin, end


class Job:
    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", self.local_rank)
            if not dist.is_initialized():
                dist.init_process_group(backend, **kw)
            self.dist = dist
            self.backend = backend

    def _tensor(self, v):
        import torch
        dev = "cuda" if (self.dist and self.backend == "nccl") else "cpu"
        return torch.tensor([float(v)], dtype=torch.float64, device=dev)

    def barrier(self):
        import torch
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def reduce_max(self, v):
        if not self.dist:
            return float(v)
        t = self._tensor(v)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(self, v):
        if not self.dist:
            return float(v)
        t = self._tensor(v)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist and self.dist.is_initialized():
            self.dist.destroy_process_group()

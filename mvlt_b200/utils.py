"""Small logging / distributed helpers with the interface the reference's engine uses (libs/utils.py:18-161:
``SmoothedValue`` / ``MetricLogger`` with ``update``, ``meters``, ``log_every``, ``synchronize_between_processes``)."""
from __future__ import annotations

import time
from collections import defaultdict, deque

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


class SmoothedValue:
    def __init__(self, window_size=20, fmt=None):
        self.deque = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        """libs/utils.py:36-47: sum count/total over ranks (two float64 per meter)."""
        if not is_dist():
            return
        dev = "cuda" if torch.cuda.is_available() and dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device=dev)
        dist.barrier()
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), float(t[1].item())

    @property
    def median(self):
        d = sorted(self.deque)
        return d[len(d) // 2] if d else 0.0

    @property
    def avg(self):
        return sum(self.deque) / max(len(self.deque), 1)

    @property
    def global_avg(self):
        return self.total / max(self.count, 1)

    @property
    def value(self):
        return self.deque[-1] if self.deque else 0.0

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, value=self.value)


class MetricLogger:
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, n=1, **kwargs):
        for k, v in kwargs.items():
            if isinstance(v, torch.Tensor):
                v = v.item()
            self.meters[k].update(float(v), n=n)

    def __getattr__(self, attr):
        if attr in self.__dict__.get("meters", {}):
            return self.meters[attr]
        raise AttributeError(attr)

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def synchronize_between_processes(self):
        for m in self.meters.values():
            m.synchronize_between_processes()

    def __str__(self):
        return self.delimiter.join(f"{k}: {m}" for k, m in self.meters.items())

    def log_every(self, iterable, print_freq, header=""):
        t0 = time.time()
        n = len(iterable) if hasattr(iterable, "__len__") else None
        for i, obj in enumerate(iterable):
            yield obj
            if print_freq and (i % print_freq == 0 or (n is not None and i == n - 1)):
                if not is_dist() or dist.get_rank() == 0:
                    print(f"{header} [{i}{'/' + str(n) if n else ''}] {self}  elapsed {time.time() - t0:.1f}s")


def load_file(path, trust_pickle: bool = False):
    """``torch.load`` restricted to tensors / plain containers (``weights_only=True``). The reference's full checkpoints
    (main_vl.py:327-346) also pickle the argparse ``Namespace``: those need ``trust_pickle=True``, an explicit opt-in to
    unpickling arbitrary objects from the file."""
    import argparse
    try:
        with torch.serialization.safe_globals([argparse.Namespace]):
            return torch.load(path, map_location="cpu", weights_only=True)
    except Exception as ex:
        if not trust_pickle:
            raise RuntimeError(f"{path}: not loadable with weights_only=True ({type(ex).__name__}: {str(ex)[:200]}); pass "
                               "trust_pickle=True only for checkpoints you trust") from ex
        return torch.load(path, map_location="cpu", weights_only=False)


def load_checkpoint(model, checkpoint, optimizer=None, lr_scheduler=None, loss_scaler=None, strict=False,
                    trust_pickle: bool = False):
    """Checkpoint compatibility with the reference's files. ``checkpoint``: a path or an already loaded
    object in any of the forms the reference writes / reads:
      * ``{'model': state_dict, 'optimizer': ..., 'lr_scheduler': ..., 'epoch': ..., 'scaler': ...}`` (main_vl.py:327-346,
        utils.save_on_master) or a bare ``state_dict`` (``checkpoint_retrieval.pth`` / ``checkpoint_recognition.pth``);
      * keys optionally prefixed with ``module.`` (saved from a DistributedDataParallel wrapper);
      * ImageNet ``head.*`` / ``head_dist.*`` classifier rows of another shape are dropped (main_vl.py:283-288).
    Parameter names and shapes are the reference's own (Appendix A), so nothing is renamed. Returns
    ``(missing_keys, unexpected_keys, next_epoch)``; optimizer / scheduler / scaler states are restored when both the
    object and its entry are present (torch.optim.AdamW and mvlt_b200.optim.AdamW share the state layout)."""
    net = model.module if hasattr(model, "module") else model
    ck = load_file(checkpoint, trust_pickle) if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__") \
        else checkpoint
    sd = ck["model"] if isinstance(ck, dict) and "model" in ck else ck
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
    own = net.state_dict()
    for k in ("head.weight", "head.bias", "head_dist.weight", "head_dist.bias"):
        if k in sd and (k not in own or sd[k].shape != own[k].shape):
            del sd[k]
    res = net.load_state_dict(sd, strict=strict)
    next_epoch = None
    if isinstance(ck, dict):
        if optimizer is not None and "optimizer" in ck:
            optimizer.load_state_dict(ck["optimizer"])
        if lr_scheduler is not None and "lr_scheduler" in ck:
            lr_scheduler.load_state_dict(ck["lr_scheduler"])
        if loss_scaler is not None and "scaler" in ck:
            loss_scaler.load_state_dict(ck["scaler"])
        if "epoch" in ck:
            next_epoch = int(ck["epoch"]) + 1
    return list(res.missing_keys), list(res.unexpected_keys), next_epoch

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

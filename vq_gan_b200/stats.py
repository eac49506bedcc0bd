"""Device-side training statistics without per-step host syncs.

The reference syncs three times per step for logging (`quantizer.py:106-107,147`, consumed at
`train_vqgan.py:303-315`).  With `VectorQuantizer(..., lazy_stats=True)` the losses stay on the
device; this accumulator sums them and the usage histogram on the device and copies ONE small
tensor to the host when `flush()` is called (e.g. every N steps).
"""
from typing import Dict, Optional

import torch


class DeviceStats:
    def __init__(self, num_embeddings: int, device):
        self.num_embeddings = num_embeddings
        self.usage = torch.zeros(num_embeddings, dtype=torch.int64, device=device)
        self.sums = torch.zeros(3, dtype=torch.float64, device=device)  # vq_loss, mse, steps

    @torch.no_grad()
    def update(self, loss_dict: Dict[str, torch.Tensor], usage: Optional[torch.Tensor] = None) -> None:
        """loss_dict from a lazy_stats quantizer (0-dim device tensors); usage int64[K] (optional)."""
        self.sums[0] += loss_dict["vq_loss"].detach().double()
        self.sums[1] += torch.as_tensor(loss_dict["codebook_loss"], device=self.sums.device).double()
        self.sums[2] += 1
        if usage is not None:
            self.usage += usage

    @torch.no_grad()
    def flush(self) -> Dict[str, float]:
        """One device->host copy; resets the accumulators."""
        used = (self.usage > 0).sum().double().reshape(1)
        host = torch.cat([self.sums, used, self.usage.sum().double().reshape(1)]).cpu()
        steps = max(float(host[2]), 1.0)
        out = {"vq_loss": float(host[0]) / steps, "codebook_loss": float(host[1]) / steps,
               "commitment_loss": float(host[1]) / steps, "steps": int(host[2]),
               "codebook_usage_ratio": float(host[3]) / self.num_embeddings, "tokens": int(host[4])}
        self.usage.zero_()
        self.sums.zero_()
        return out

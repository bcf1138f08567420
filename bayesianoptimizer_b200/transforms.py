"""Output transforms around the engine (elementwise torch plumbing, stays on the tensors' device).

``LogStandardize`` is the log + standardise transform the reference's exact 8-output model trains on
(optimization/Bayesian6.py:427-443, 462-468) and its lognormal back-transform ``exp(mu + var/2) - shift``
(:631-633, :703-707).  ``Standardize`` (botorch's default outcome transform) lives in
``BayesianOptimizer._model_targets``.
"""
from __future__ import annotations

import torch


class LogStandardize:
    def __init__(self, shift: float, mean: torch.Tensor, std: torch.Tensor):
        self.shift, self.mean, self.std = float(shift), mean, std

    @classmethod
    def fit(cls, Y: torch.Tensor) -> "LogStandardize":
        Y = torch.as_tensor(Y, dtype=torch.float64)
        Y = Y.reshape(Y.shape[0], -1)
        eps = max(1e-12, float(Y.abs().max()) * 1e-6) if Y.numel() else 1e-6          # _compute_safe_epsilon
        ymin = float(Y.min())
        shift = (-ymin + eps) if ymin <= 0.0 else eps
        lg = torch.log(Y + shift)
        std = lg.std(dim=0, keepdim=True)
        std = torch.where(std < 1e-12, torch.full_like(std, 1e-12), std)
        return cls(shift, lg.mean(dim=0, keepdim=True), std)

    def forward(self, Y: torch.Tensor) -> torch.Tensor:
        Y = torch.as_tensor(Y, dtype=torch.float64, device=self.mean.device)
        Y = Y.reshape(Y.shape[0], -1)
        return torch.nan_to_num((torch.log(Y + self.shift) - self.mean) / self.std, nan=0.0)

    def inverse_mean(self, mean_std: torch.Tensor, var_std: torch.Tensor) -> torch.Tensor:
        """Mean of the lognormal predictive in raw units; ``var_std`` (N,) or (N,1) broadcasts over the outputs."""
        var_std = torch.as_tensor(var_std, dtype=torch.float64, device=self.mean.device)
        if var_std.ndim == 1:
            var_std = var_std.unsqueeze(-1)
        log_mean = torch.as_tensor(mean_std, dtype=torch.float64, device=self.mean.device) * self.std + self.mean
        log_var = var_std * self.std ** 2
        return torch.exp(log_mean + 0.5 * log_var) - self.shift

"""Drop-in PPO buffer: `RMABPPO_Buffer` of robust_rmab/algos/rmabppo/agent_oracle.py:24-161 with the per-arm
Python loops of `finish_path` (GAE, rewards-to-go, cost-to-go) and `get` (advantage normalisation) replaced by
the segmented-scan kernels (csrc/gae.cu).

Kept from the reference: constructor signature, `store(obs, act, rew, cost, val, q, lamb, logp)`,
`finish_path(last_vals, last_costs)`, `get() -> dict` with keys obs, act, ret, adv, logp, qs, oha, ohs, costs,
lambdas as float32 CPU tensors of the reference's shapes (`[T,N]`, `[T,N,obs_dim]`, `[T]`), and the
`assert self.ptr < self.max_size` / `== max_size` checks.
New: `n_envs=E` (stored rows are `[N,E]`; returned arrays gain a trailing env axis when E > 1) and
`get_device()` which returns the `[N,T,E]` device tensors without a host copy.
"""
import numpy as np
import torch

from . import ops

_F32 = torch.float32


class RMABPPO_Buffer:

    def __init__(self, obs_dim, act_dim, N, act_type, size, one_hot_encode=True, gamma=0.99, lam_OTHER=0.95, n_envs=1,
                 device="cuda"):
        self.N, self.obs_dim, self.act_dim, self.act_type = int(N), int(obs_dim), int(act_dim), act_type
        self.one_hot_encode = one_hot_encode
        self.gamma, self.lam_OTHER = gamma, lam_OTHER
        self.ptr, self.path_start_idx, self.max_size = 0, 0, int(size)
        self.E = int(n_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RMABPPO_Buffer computes on CUDA only (no CPU fallback)")
        N, T, E = self.N, self.max_size, self.E
        z = lambda: torch.zeros((N, T, E), dtype=_F32, device=self.device)
        self.obs_buf, self.act_buf, self.rew_buf, self.cost_buf = z(), z(), z(), z()
        self.val_buf, self.q_buf, self.logp_buf = z(), z(), z()
        self.adv_buf, self.ret_buf = z(), z()
        self.lamb_buf = torch.zeros((T, E), dtype=_F32, device=self.device)
        self.cdcost_buf = torch.zeros((T, E), dtype=_F32, device=self.device)

    def _row(self, x):
        if isinstance(x, torch.Tensor):
            t = x.detach().to(self.device, _F32)
        else:
            t = torch.as_tensor(np.asarray(x, dtype=np.float32), device=self.device)
        return t.reshape(self.N, self.E)

    def store(self, obs, act, rew, cost, val, q, lamb, logp):
        assert self.ptr < self.max_size     # buffer has to have room so you can store
        p = self.ptr
        self.obs_buf[:, p], self.act_buf[:, p] = self._row(obs), self._row(act)
        self.rew_buf[:, p], self.cost_buf[:, p] = self._row(rew), self._row(cost)
        self.val_buf[:, p], self.q_buf[:, p], self.logp_buf[:, p] = self._row(val), self._row(q), self._row(logp)
        lam = lamb.detach().to(self.device, _F32) if isinstance(lamb, torch.Tensor) else \
            torch.as_tensor(np.asarray(lamb, dtype=np.float32), device=self.device)
        self.lamb_buf[p] = lam.reshape(-1).expand(self.E) if lam.numel() == 1 else lam.reshape(self.E)
        self.ptr += 1

    def finish_path(self, last_vals=0, last_costs=0):
        """GAE-lambda advantages, rewards-to-go, q and the arm-summed discounted cost-to-go of the path
        [path_start_idx, ptr) (agent_oracle.py:90-139).  The reference charges each step its own stored lambda
        (:115-120).  Its training loop predicts lambda once per epoch (:497), so a path normally has one lambda per
        env and the kernel applies it; a caller that stored per-step lambdas gets them folded into the reward
        (r_t - lambda_t c_t in fp32, as numpy does there) before the same scan."""
        a, b = self.path_start_idx, self.ptr
        if b == a:
            return
        if np.isscalar(last_vals):
            lv = torch.full((self.N, self.E), float(last_vals), dtype=_F32, device=self.device)
        else:
            lv = self._row(last_vals).contiguous()
        sl = lambda t: t[:, a:b].contiguous()
        lamb = self.lamb_buf[a].contiguous()
        rew, cost = sl(self.rew_buf), sl(self.cost_buf)
        if b - a > 1 and not bool((self.lamb_buf[a:b] == self.lamb_buf[a]).all()):
            rew = (rew - self.lamb_buf[a:b][None] * cost).contiguous()      # per-step lambdas (:115-120)
            cost, lamb = torch.zeros_like(cost), torch.zeros_like(lamb)
        adv, ret, q, _ = ops.gae_explicit(rew, cost, sl(self.val_buf), lv, lamb, self.gamma,
                                          self.lam_OTHER, normalize=False, want_q=True)
        self.adv_buf[:, a:b], self.ret_buf[:, a:b], self.q_buf[:, a:b] = adv, ret, q
        # cost-to-go of the arm-summed cost (:117, :139); costs are C[a] with small integer action indices
        cost = sl(self.cost_buf)
        levels = torch.unique(cost)
        idx = torch.bucketize(cost, levels).to(torch.uint8)
        self.cdcost_buf[a:b] = ops.cost_togo(idx.contiguous(), levels.to(_F32).contiguous(), self.gamma)
        self.path_start_idx = self.ptr

    def _normalised_adv(self):
        mean = self.adv_buf.mean(dim=(1, 2), keepdim=True)
        std = (self.adv_buf - mean).pow(2).mean(dim=(1, 2), keepdim=True).sqrt()   # population std (mpi_tools.py:87-92)
        return (self.adv_buf - mean) / std

    def get_device(self):
        assert self.ptr == self.max_size    # buffer has to be full before you can get
        self.ptr, self.path_start_idx = 0, 0
        self.adv_buf = self._normalised_adv()
        return dict(obs=self.obs_buf, act=self.act_buf, ret=self.ret_buf, adv=self.adv_buf, logp=self.logp_buf,
                    qs=self.q_buf, costs=self.cdcost_buf, lambdas=self.lamb_buf)

    def get(self):
        d = self.get_device()
        tn = lambda t: t.permute(1, 0, 2)          # [N,T,E] -> [T,N,E]
        out = {k: tn(d[k]) for k in ("obs", "act", "ret", "adv", "logp", "qs")}
        out["costs"], out["lambdas"] = d["costs"], d["lambdas"]
        obs_i, act_i = out["obs"].long(), out["act"].long()
        ohs = torch.zeros(obs_i.shape + (self.obs_dim,), dtype=_F32, device=self.device)
        if self.one_hot_encode:
            ohs.scatter_(-1, obs_i[..., None], 1.0)
        oha = torch.zeros(act_i.shape + (self.act_dim,), dtype=_F32, device=self.device)
        oha.scatter_(-1, act_i[..., None], 1.0)
        out["ohs"], out["oha"] = ohs, oha
        res = {}
        for k, v in out.items():
            v = v.detach().to("cpu", _F32)
            if self.E == 1:   # the reference's shapes: [T,N], [T,N,dim], [T]
                v = v[:, :, 0] if k in ("obs", "act", "ret", "adv", "logp", "qs") else \
                    (v[:, :, 0, :] if k in ("ohs", "oha") else v[:, 0])
            res[k] = v.contiguous()
        return res

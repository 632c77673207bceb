"""The dense nature nets of MA-RMABPPO -- Gaussian actor N -> h -> h -> D (ma_rmabppo_core.py:84-113) and critic
2N -> h -> h -> 1 (:116-123) -- as packed device parameters driven through the rmab_nat_* kernels (csrc/nature.cu):
forward for the rollout (:274-331) and the PPO / MSE Adam steps of nature_oracle.py:359-386, :421-428, :558-576.

Packed layout `[W1T (Nin x H) | b1 | W2 | b2 | W3 (OUT x H) | b3 | log_std]`, W1 transposed: the rows that belong to an
arm are contiguous, so an arm shard is a set of contiguous slices.  Sharded training: a rank evaluates
layer 1 over ITS arms (partial sums `[M,H]`, one all-reduce), the actor's output layer over ITS D/N rows per arm (log-prob
partial `[M]` and dLoss/da2 partial `[M,H]`, one all-reduce each), updates its slices with Adam and finally all ranks
exchange the slices.  The replicated middle layer sees identical gradients on every rank.  torch is only the memory and
the collectives; the torch modules of the drop-in (`ac.pi_nature`, `ac.v_nature`) are kept in sync with the packed form.
"""
import torch

from . import ops

_F32 = torch.float32


class PackedDenseNet:
    def __init__(self, seq, n_in, H, out_dim, dev, log_std=None):
        """seq: the torch Sequential (Linear, act, Linear, act, Linear) the drop-in object exposes."""
        self.seq, self.log_std_param = seq, log_std
        self.Nin, self.H, self.OUT, self.dev = int(n_in), int(H), int(out_dim), dev
        H, Nin, OUT = self.H, self.Nin, self.OUT
        self.oW1, self.ob1 = 0, Nin * H
        self.oW2 = self.ob1 + H
        self.ob2 = self.oW2 + H * H
        self.oW3 = self.ob2 + H
        self.ob3 = self.oW3 + OUT * H
        self.ols = self.ob3 + OUT
        self.P = self.ols + (OUT if log_std is not None else 0)
        self.p = torch.empty(self.P, dtype=_F32, device=dev)
        self.g = torch.zeros_like(self.p)
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.step = 0
        self.pack()

    def _linears(self):
        return [m for m in self.seq if isinstance(m, torch.nn.Linear)]

    @torch.no_grad()
    def pack(self):
        """module -> packed (the tests and callers may have written the module's parameters in place)."""
        l1, l2, l3 = self._linears()
        p, H = self.p, self.H
        p[self.oW1:self.ob1].copy_(l1.weight.detach().t().reshape(-1))
        p[self.ob1:self.oW2].copy_(l1.bias.detach())
        p[self.oW2:self.ob2].copy_(l2.weight.detach().reshape(-1))
        p[self.ob2:self.oW3].copy_(l2.bias.detach())
        p[self.oW3:self.ob3].copy_(l3.weight.detach().reshape(-1))
        p[self.ob3:self.ols].copy_(l3.bias.detach())
        if self.log_std_param is not None:
            p[self.ols:].copy_(self.log_std_param.detach())

    @torch.no_grad()
    def unpack(self):
        """packed -> module."""
        l1, l2, l3 = self._linears()
        p, H = self.p, self.H
        l1.weight.copy_(p[self.oW1:self.ob1].view(self.Nin, H).t())
        l1.bias.copy_(p[self.ob1:self.oW2])
        l2.weight.copy_(p[self.oW2:self.ob2].view(H, H))
        l2.bias.copy_(p[self.ob2:self.oW3])
        l3.weight.copy_(p[self.oW3:self.ob3].view(self.OUT, H))
        l3.bias.copy_(p[self.ob3:self.ols])
        if self.log_std_param is not None:
            self.log_std_param.copy_(p[self.ols:])

    # views into a packed vector (parameters or gradients)
    def W1T(self, buf, row0, rows):
        return buf[self.oW1 + row0 * self.H:self.oW1 + (row0 + rows) * self.H]

    def b1(self, buf):
        return buf[self.ob1:self.oW2]

    def W2(self, buf):
        return buf[self.oW2:self.ob2]

    def b2(self, buf):
        return buf[self.ob2:self.oW3]

    def W3(self, buf, row0=0, rows=None):
        rows = self.OUT if rows is None else rows
        return buf[self.oW3 + row0 * self.H:self.oW3 + (row0 + rows) * self.H]

    def b3(self, buf, row0=0, rows=None):
        rows = self.OUT if rows is None else rows
        return buf[self.ob3 + row0:self.ob3 + row0 + rows]

    def ls(self, buf, row0=0, rows=None):
        rows = self.OUT if rows is None else rows
        return buf[self.ols + row0:self.ols + row0 + rows]

    def adam(self, slices, lr):
        """One Adam step (torch.optim.Adam defaults) on the given (start, end) slices of the packed vector."""
        self.step += 1
        for a, b in slices:
            if b > a:
                ops.adam_step(self.p[a:b], self.g[a:b], self.m[a:b], self.v[a:b], lr, self.step)


class NatureNets:
    """Actor + critic in packed form, with the scratch buffers of one sample count M."""

    def __init__(self, ac, arm0=0, n_loc=None):
        self.ac = ac
        self.N, self.D, self.H = ac.N, ac.act_dim_nature, ac.hidden_sizes[0]
        self.dev = ac.device_
        self.k = self.D // self.N                       # nature outputs per arm
        self.arm0 = int(arm0)
        self.n_loc = self.N if n_loc is None else int(n_loc)
        self.actor = PackedDenseNet(ac.pi_nature.mu_net, self.N, self.H, self.D, self.dev, log_std=ac.pi_nature.log_std)
        self.critic = PackedDenseNet(ac.v_nature.v_net, 2 * self.N, self.H, 1, self.dev)
        self._scratch = {}

    def pack(self):
        self.actor.pack()
        self.critic.pack()

    def unpack(self):
        self.actor.unpack()
        self.critic.unpack()

    def _buf(self, M):
        b = self._scratch.get(M)
        if b is None:
            H, dev = self.H, self.dev
            z = lambda *s: torch.empty(s, dtype=_F32, device=dev)
            b = dict(z1=z(M, H), a1=z(M, H), a2=z(M, H), da2=z(M, H), d1=z(M, H), d2=z(M, H), logp=z(M), g=z(M), v=z(M),
                     ws=ops.nat_ws(M, H, max(self.N, 1), max(self.D, 1), dev))
            self._scratch[M] = b
        return b

    # ---- forward over ALL arms (rollout; every rank evaluates the same thing) ----------------------------------
    def actor_mu(self, s_full, scale):
        """s_full `[N,E]` u8 -> mu `[E,D]`."""
        E = s_full.shape[1]
        b, a, H = self._buf(E), self.actor, self.H
        ops.nat_l1_fwd(E, E, H, s_full, scale, a.W1T(a.p, 0, self.N), b["z1"], b["ws"])
        ops.nat_mid_fwd(E, H, b["z1"], a.b1(a.p), a.W2(a.p), a.b2(a.p), b["a1"], b["a2"])
        mu = torch.empty((E, self.D), dtype=_F32, device=self.dev)
        return ops.nat_mu(E, H, self.D, b["a2"], a.W3(a.p), a.b3(a.p), mu)

    def critic_values(self, s, a_agent, scale, T_rows, E, allreduce=None):
        """Nature critic over [obs, a_agent]: s `[n_loc, rows >= T_rows, E]`, a_agent `[n_loc, T_rows, E]` u8 (this rank's
        arms) -> v `[T_rows * E]`; `allreduce` sums the layer-1 partials when the arms are sharded."""
        M = T_rows * E
        b, c, H = self._buf(M), self.critic, self.H
        ops.nat_l1_fwd(M, E, H, s, scale, c.W1T(c.p, self.arm0, self.n_loc), b["z1"], b["ws"], x1=a_agent,
                       W1=c.W1T(c.p, self.N + self.arm0, self.n_loc))
        if allreduce is not None:
            allreduce(b["z1"])
        ops.nat_mid_fwd(M, H, b["z1"], c.b1(c.p), c.W2(c.p), c.b2(c.p), b["a1"], b["a2"])
        return ops.nat_v_out(M, H, b["a2"], c.W3(c.p), c.b3(c.p), v=torch.empty((M,), dtype=_F32, device=self.dev))

    # ---- updates (sharded by arm) --------------------------------------------------------------------------------
    def actor_step(self, s_traj, T, E, scale, act, logp_old, adv, clip, lr, allreduce=None):
        """One Adam step of the nature policy (nature_oracle.py:359-386, :558-567).  s_traj `[n_loc, >= T, E]` u8, act `[T*E, D]`
        raw nature actions, logp_old / adv `[T*E]`."""
        M, H, a, k = T * E, self.H, self.actor, self.k
        b = self._buf(M)
        d0, Dl = self.arm0 * k, self.n_loc * k
        ops.nat_l1_fwd(M, E, H, s_traj, scale, a.W1T(a.p, self.arm0, self.n_loc), b["z1"], b["ws"])
        if allreduce is not None:
            allreduce(b["z1"])
        ops.nat_mid_fwd(M, H, b["z1"], a.b1(a.p), a.W2(a.p), a.b2(a.p), b["a1"], b["a2"])
        act_loc = act[:, d0:] if Dl else act
        ops.nat_logp(M, H, Dl, b["a2"], a.W3(a.p, d0, Dl), a.b3(a.p, d0, Dl), a.ls(a.p, d0, Dl), act_loc, b["logp"], b["ws"])
        if allreduce is not None:
            allreduce(b["logp"])
        ops.nat_pi_coef(M, b["logp"], logp_old, adv, clip, b["g"])
        ops.nat_out_bwd(M, H, Dl, b["a2"], a.W3(a.p, d0, Dl), a.b3(a.p, d0, Dl), a.ls(a.p, d0, Dl), act_loc, b["g"],
                        a.W3(a.g, d0, Dl), a.b3(a.g, d0, Dl), a.ls(a.g, d0, Dl), b["da2"], b["ws"])
        if allreduce is not None:
            allreduce(b["da2"])
        ops.nat_mid_bwd(M, H, b["da2"], b["a1"], b["a2"], a.W2(a.p), b["d1"], b["d2"], a.W2(a.g), a.b2(a.g), a.b1(a.g), b["ws"])
        ops.nat_l1_bwd(M, E, H, s_traj, scale, b["d1"], a.W1T(a.g, self.arm0, self.n_loc), b["ws"])
        o = a
        a.adam([(o.oW1 + self.arm0 * H, o.oW1 + (self.arm0 + self.n_loc) * H), (o.ob1, o.oW3),
                (o.oW3 + d0 * H, o.oW3 + (d0 + Dl) * H), (o.ob3 + d0, o.ob3 + d0 + Dl), (o.ols + d0, o.ols + d0 + Dl)], lr)

    def critic_step(self, s_traj, a_agent, T, E, scale, ret, lr, allreduce=None):
        """One Adam step of the nature critic (nature_oracle.py:421-428, :570-576)."""
        M, H, c = T * E, self.H, self.critic
        b = self._buf(M)
        ops.nat_l1_fwd(M, E, H, s_traj, scale, c.W1T(c.p, self.arm0, self.n_loc), b["z1"], b["ws"], x1=a_agent,
                       W1=c.W1T(c.p, self.N + self.arm0, self.n_loc))
        if allreduce is not None:
            allreduce(b["z1"])
        ops.nat_mid_fwd(M, H, b["z1"], c.b1(c.p), c.W2(c.p), c.b2(c.p), b["a1"], b["a2"])
        ops.nat_v_out(M, H, b["a2"], c.W3(c.p), c.b3(c.p), ret=ret, da2=b["da2"], dW3=c.W3(c.g), db3=c.b3(c.g), ws=b["ws"])
        ops.nat_mid_bwd(M, H, b["da2"], b["a1"], b["a2"], c.W2(c.p), b["d1"], b["d2"], c.W2(c.g), c.b2(c.g), c.b1(c.g), b["ws"])
        ops.nat_l1_bwd(M, E, H, s_traj, scale, b["d1"], c.W1T(c.g, self.arm0, self.n_loc), b["ws"], x1=a_agent,
                       dW1=c.W1T(c.g, self.N + self.arm0, self.n_loc))
        o = c
        c.adam([(o.oW1 + self.arm0 * H, o.oW1 + (self.arm0 + self.n_loc) * H),
                (o.oW1 + (self.N + self.arm0) * H, o.oW1 + (self.N + self.arm0 + self.n_loc) * H), (o.ob1, o.P)], lr)

    def exchange_shards(self, allreduce):
        """After sharded updates: every rank contributes the slices it owns (arm rows of W1T, W3, b3, log_std); the
        replicated middle is already identical everywhere."""
        H, k = self.H, self.k
        a0, a1 = self.arm0, self.arm0 + self.n_loc
        for net, spans in ((self.actor, [(self.actor.oW1 + a0 * H, self.actor.oW1 + a1 * H),
                                         (self.actor.oW3 + a0 * k * H, self.actor.oW3 + a1 * k * H),
                                         (self.actor.ob3 + a0 * k, self.actor.ob3 + a1 * k),
                                         (self.actor.ols + a0 * k, self.actor.ols + a1 * k)]),
                           (self.critic, [(self.critic.oW1 + a0 * H, self.critic.oW1 + a1 * H),
                                          (self.critic.oW1 + (self.N + a0) * H, self.critic.oW1 + (self.N + a1) * H)])):
            tmp = torch.zeros_like(net.p)
            for s_, e_ in spans:
                tmp[s_:e_] = net.p[s_:e_]
            allreduce(tmp)
            # positions some rank owns: W1T / W3 / b3 / log_std blocks; everything else keeps the local (identical) value
            mask = torch.zeros(net.P, dtype=torch.bool, device=self.dev)
            mask[net.oW1:net.ob1] = True
            if net is self.actor:
                mask[net.oW3:net.P] = True
            net.p.copy_(torch.where(mask, tmp, net.p))

"""Policy sampler with the reference's call shapes (models/controller.py:9-145):
`Controller(cfg)(M) -> (policies int64 [M, Q*L*2], op_probs [NUM_OPS], mag_probs [NUM_MAGS], log_probs [M],
entropies [M])`, `controller.evaluate(policies, M) -> log_probs [M]`.  Policies are laid out
(op, mag) x L x Q like the reference (controller.py:110, decoded by data/policy.py:93); the hidden
state restarts for every sub-policy (controller.py:81).  ~50 kFLOP once per epoch: plain torch."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..data.basic import augment_list


class Controller(nn.Module):
    def __init__(self, cfg, n_subpolicies=5, embedding_dim=32, hidden_dim=100):
        super().__init__()
        c = cfg.CONTROLLER
        self.L, self.T, self.C, self.NUM_MAGS, self.Q = c.L, c.T, c.C, c.NUM_MAGS, n_subpolicies
        excluded = len(c.EXCLUDE_OPS) if len(c.EXCLUDE_OPS) > 0 else c.EXCLUDE_OPS_NUM
        self.NUM_OPS = len(augment_list()) - excluded
        self.embedding_dim, self.hidden_dim = embedding_dim, hidden_dim
        self.embedding = nn.Embedding(self.NUM_OPS + self.NUM_MAGS, embedding_dim)
        self.lstm = nn.LSTMCell(embedding_dim, hidden_dim)
        self.outop = nn.Linear(hidden_dim, self.NUM_OPS)
        self.outmag = nn.Linear(hidden_dim, self.NUM_MAGS)
        for p in self.parameters():                      # controller.py:31-36
            p.data.uniform_(-0.1, 0.1)
        self.outop.bias.data.zero_()
        self.outmag.bias.data.zero_()

    def _start(self, m):
        dev = self.embedding.weight.device
        return (torch.zeros(m, self.embedding_dim, device=dev), torch.zeros(m, self.hidden_dim, device=dev),
                torch.zeros(m, self.hidden_dim, device=dev))

    def _log_softmax(self, logits):
        return F.log_softmax(self.C * torch.tanh(logits) / self.T, dim=-1)

    def _walk(self, m, choose):
        """One pass over Q sub-policies x L (op, mag) decisions; `choose(step, log_prob)` returns the action."""
        for i in range(self.Q):
            inp, hx, cx = self._start(m)
            for j in range(self.L):
                for head, offset, kind in ((self.outop, 0, "op"), (self.outmag, self.NUM_OPS, "mag")):
                    hx, cx = self.lstm(inp, (hx, cx))
                    logp = self._log_softmax(head(hx))
                    action = choose((i * self.L + j) * 2 + (kind == "mag"), kind, logp)
                    inp = self.embedding(offset + action)

    def forward(self, batch_size=1):
        return self.sample(batch_size)

    def sample(self, batch_size=1):
        acts, logps, ents, opp, magp = [], [], [], [], []

        def choose(step, kind, logp):
            probs = logp.exp()
            action = probs.multinomial(num_samples=1)[:, 0]
            acts.append(action)
            logps.append(logp.gather(1, action[:, None])[:, 0])
            ents.append(-(logp * probs).sum(1))
            (opp if kind == "op" else magp).append(probs)
            return action
        self._walk(batch_size, choose)
        policies = torch.stack(acts, dim=-1)
        op_probs = torch.stack(opp, dim=-1).permute(0, 2, 1).reshape(-1, self.NUM_OPS)
        mag_probs = torch.stack(magp, dim=-1).permute(0, 2, 1).reshape(-1, self.NUM_MAGS)
        return (policies, op_probs.mean(0), mag_probs.mean(0), torch.stack(logps, -1).sum(-1),
                torch.stack(ents, -1).sum(-1))

    def evaluate(self, policies, batch_size):
        logps = []

        def choose(step, kind, logp):
            action = policies[:, step].long()
            logps.append(logp.gather(1, action[:, None])[:, 0])
            return action
        self._walk(batch_size, choose)
        return torch.stack(logps, -1).sum(-1)


def evaluate_with_entropy(controller, policies, batch_size):
    """(sum_t log pi(a_t), sum_t H_t) [M] WITH an autograd graph, through the torch mirror of the walk (what the
    reference's `controller(M)` returns for the REINFORCE loss, losses.py:96-114).  Used when the sampled
    log-probabilities carry no graph (FusedController.sample)."""
    logps, ents = [], []

    def choose(step, kind, logp):
        action = policies[:, step].long()
        logps.append(logp.gather(1, action[:, None])[:, 0])
        ents.append(-(logp * logp.exp()).sum(1))
        return action
    Controller._walk(controller, batch_size, choose)
    return torch.stack(logps, -1).sum(-1), torch.stack(ents, -1).sum(-1)


class _WalkEvaluate(torch.autograd.Function):
    """sum_t log pi(a_t) for given policies: forward = one aadg_controller_walk launch (mode 1), backward = one
    aadg_controller_backward launch producing the gradients of all nine parameter tensors."""

    @staticmethod
    def forward(ctx, ctl, policies, *params):
        step_logp, _, _, saved, _ = ctl._walk(policies, mode=1, want_saved=True)
        ctx.ctl, ctx.policies, ctx.saved = ctl, policies, saved
        ctx.save_for_backward(*params)
        return step_logp.sum(-1)

    @staticmethod
    def backward(ctx, grad_out):
        ctl = ctx.ctl
        grads = [torch.zeros_like(p) for p in ctx.saved_tensors]
        ctl._backward(ctx.policies, ctx.saved, grad_out.contiguous().float(), grads)
        return (None, None) + tuple(grads)


class FusedController(Controller):
    """Same module, parameters, state_dict and call shapes as `Controller` (models/controller.py), but `sample` and
    `evaluate` are one kernel launch each (csrc/controller.cu) and `evaluate` back-propagates through one more, so the
    reference's PPO loop (losses.py:127-157: five evaluate / backward / Adam rounds per epoch) needs 10 launches of
    this library instead of ~1 000 small ATen ones.  Sampling is counter based (Philox keyed by `seed` and the call
    number) instead of torch's multinomial stream.  CUDA only."""

    def __init__(self, cfg, n_subpolicies=5, embedding_dim=32, hidden_dim=100, seed=0):
        super().__init__(cfg, n_subpolicies, embedding_dim, hidden_dim)
        self.seed, self.calls = int(seed), 0

    # the Philox call counter travels with the state_dict (a resumed run must not replay the same sample stream); it
    # is stored as module extra state so that the parameter / buffer key set stays the reference Controller's
    def get_extra_state(self):
        return {"calls": int(self.calls), "seed": int(self.seed)}

    def set_extra_state(self, state):
        self.calls = int(state.get("calls", 0))
        self.seed = int(state.get("seed", self.seed))

    def load_state_dict(self, state_dict, strict=True, **kw):
        """accepts a reference `Controller` checkpoint (no `_extra_state` key) as well as its own"""
        if "_extra_state" not in state_dict:
            state_dict = dict(state_dict)
            state_dict["_extra_state"] = self.get_extra_state()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _params(self):
        return [self.embedding.weight, self.lstm.weight_ih, self.lstm.weight_hh, self.lstm.bias_ih, self.lstm.bias_hh,
                self.outop.weight, self.outop.bias, self.outmag.weight, self.outmag.bias]

    @staticmethod
    def _ptr_array(tensors):
        import ctypes
        arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        return arr, ctypes.addressof(arr)

    def _check(self):
        from .. import _lib
        ps = self._params()
        if not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in ps):
            raise RuntimeError("FusedController needs contiguous float32 CUDA parameters (no CPU path)")
        return _lib, ps

    def _walk(self, policies, mode, want_saved=False, want_probs=False, batch=None):
        _lib, ps = self._check()
        dev = ps[0].device
        steps = self.Q * self.L * 2
        m = policies.shape[0] if policies is not None else batch
        if policies is None:
            policies = torch.empty((m, steps), dtype=torch.int64, device=dev)
        elif not (policies.is_cuda and policies.dtype == torch.int64 and policies.is_contiguous()
                  and policies.shape == (m, steps)):
            raise ValueError("policies must be a contiguous CUDA int64 [batch, %d] tensor" % steps)
        logp = torch.empty((m, steps), dtype=torch.float32, device=dev)
        ent = torch.empty((m, steps), dtype=torch.float32, device=dev)
        vm = max(self.NUM_OPS, self.NUM_MAGS)
        probs = torch.empty((m, steps, vm), dtype=torch.float32, device=dev) if want_probs else None
        saved = torch.empty((m, steps, 6 * self.hidden_dim), dtype=torch.float32, device=dev) if want_saved else None
        arr, addr = self._ptr_array([p.detach() for p in ps])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().aadg_controller_walk(
                addr, self.NUM_OPS, self.NUM_MAGS, self.Q, self.L, self.embedding_dim, self.hidden_dim, float(self.C),
                float(self.T), m, mode, self.seed, self.calls, policies.data_ptr(), logp.data_ptr(), ent.data_ptr(),
                None if probs is None else probs.data_ptr(), None if saved is None else saved.data_ptr(),
                _lib.stream_ptr()))
        del arr
        return logp, ent, probs, saved, policies

    def _backward(self, policies, saved, grad, grads):
        _lib, ps = self._check()
        parr, paddr = self._ptr_array([p.detach() for p in ps])
        garr, gaddr = self._ptr_array(grads)
        with torch.cuda.device(ps[0].device):
            _lib.check(_lib.lib().aadg_controller_backward(
                paddr, self.NUM_OPS, self.NUM_MAGS, self.Q, self.L, self.embedding_dim, self.hidden_dim, float(self.C),
                float(self.T), policies.shape[0], policies.data_ptr(), saved.data_ptr(), grad.data_ptr(), gaddr,
                _lib.stream_ptr()))
        del parr, garr

    def sample(self, batch_size=1):
        """(policies int64 [M, Q*L*2], op_probs [NUM_OPS], mag_probs [NUM_MAGS], log_probs [M], entropies [M]):
        the reference's return tuple (controller.py:118-119); log_probs carries no graph (the PPO loss detaches it)."""
        logp, ent, probs, _, policies = self._walk(None, mode=0, want_probs=True, batch=batch_size)
        self.calls += 1
        op_probs = probs[:, 0::2, :self.NUM_OPS].reshape(-1, self.NUM_OPS).mean(0)
        mag_probs = probs[:, 1::2, :self.NUM_MAGS].reshape(-1, self.NUM_MAGS).mean(0)
        return policies, op_probs, mag_probs, logp.sum(-1), ent.sum(-1)

    def evaluate(self, policies, batch_size):
        policies = policies.long().contiguous()
        return _WalkEvaluate.apply(self, policies, *self._params())

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

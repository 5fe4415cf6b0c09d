"""Score-provider adapters: the operands of a model's `predict` as (user_vecs, item_vecs, bias).

The fused kernels score `user_vecs[B, d] @ item_vecs[I, d].T (+ bias[I])` tile by tile without ever
materialising the `[B, I]` block the reference's `predict` returns (and copies to the host,
`...cpu().detach().numpy()`, e.g. LightGCN.py:214-216).  Every scorer below is an exact rewrite of the
reference model's scoring expression into that form, or -- where stated -- a strictly monotone
transform of it per user, which leaves every rank list and therefore every metric unchanged.

    model.eval_embeddings = adapters.dot_product(U, I, b).bind()        # or
    evaluator.evaluate(adapters.dot_product(U, I, b))                   # the scorer is itself a model

| reference model(s)                                   | scoring expression (file:line)                       | adapter |
|---|---|---|
| BPRMF                                                | u.i + b_i            (BPRMF.py:84-88)                | dot_product(U, I, b) |
| LightGCN, LayerGCN, LightGCL, SLMRec, SGL-style GCNs | u_final.i_final      (LightGCN.py:102-107)           | dot_product(U_final, I_final) |
| FREEDOM, LATTICE, MGCN, DENS, AOBPR                  | u.i on the propagated / learnt tables (FREEDOM.py:254-260, LATTICE.py:296-302, MGCN.py:355-361, DENS.py:315-316, AOBPR.py:94-97) | dot_product(U, I) |
| BM3                                                  | pred(u_on).pred(i_on) (BM3.py:206-210)               | dot_product(pred(U_on), pred(I_on)) |
| SLMRec                                               | sigmoid(u.i)         (SLMRec.py:366-370)             | dot_product(U, I)  (monotone: the sigmoid is dropped) |
| SASRec, SRGNN                                        | h_last.i             (SASRec.py:463, SRGNN.py:176)   | dot_product(H_last, I): one query row per evaluated user |
| GRU4Rec, GRU4RecPlus                                 | act(h.i + b_i)       (GRU4Rec.py:157-158)            | dot_product(H, I, b)  (monotone final activation dropped) |
| HGN                                                  | (u + union + sum_l e_l).W2 + b2 (HGN.py:147-163)     | summed_query([u, union, e.sum(1)], W2, b2) |
| SelfCF                                               | u_on.i_tg + u_tg.i_on (SelfCF.py:235-241)            | two_tower_sum(u_on, i_tg, u_tg, i_on) |
| FPMC                                                 | ui.iu + last.il      (FPMC.py:81-87)                 | two_tower_sum(UI, IU, LI[last], IL) |
| MultVAE, CDAE                                        | h(x_u).W^T + c       (MultVAE.py:138-141)            | decoder_layer(H, W, c) |
| CML                                                  | -||u - i||           (CML.py:152)                    | neg_euclidean(U, I)  (monotone: 2u.i - ||i||^2) |
| Pop                                                  | popularity count     (Pop.py:41-44)                  | item_scores(counts) |
| TransRec                                             | -||u + g + last - i|| + b_i (TransRec.py:86-93)      | transrec(U, g, I, b, last_items): neg_l2_plus_bias on the translated queries |
| SGAT                                                 | -||head + u - i|| + b_i     (SGAT.py:300)            | neg_l2_plus_bias(head + u, I, b) |

`neg_l2_plus_bias` is not a contraction: the square root next to a per-item bias is not a monotone image of a dot
product.  The scorer carries `score_fn = "neg_l2"`; the evaluator then runs the FP32 tile kernels with a distance
inner loop (sum of (q_k - i_k)^2, k ascending, -sqrt in the epilogue, + bias) -- still fused, nothing is materialised.

Activations: "monotone final activation dropped" holds for strictly increasing ones (linear, leaky_relu, tanh, sigmoid
up to float saturation).  GRU4Rec's `final_act` may be relu (GRU4Rec.py:274-276), which is only weakly monotone: all
negative scores collapse to 0 and tie, so a user with fewer than K positive scores gets a different tail of the rank
list than the reference (which orders the tied zeros by its heap).  For relu keep the model's own `predict`.
"""
import numpy as np

__all__ = ["EmbeddingScorer", "dot_product", "two_tower_sum", "summed_query", "decoder_layer", "neg_euclidean", "item_scores",
           "neg_l2_plus_bias", "transrec"]


def _t(x):
    import torch
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.detach()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


class EmbeddingScorer(object):
    """Holds (user table, item table, bias) and speaks both protocols: `eval_embeddings(users)` for the
    fused path and the reference's `predict(users) -> float32 ndarray [B, num_items]` (base.py:73).
    `user_index` maps user ids to rows of the user table (default: identity)."""

    def __init__(self, user_table, item_table, bias=None, user_index=None, note="dot"):
        import torch
        self.user_table, self.item_table, self.bias = _t(user_table), _t(item_table), _t(bias)
        assert self.user_table.dim() == 2 and self.item_table.dim() == 2
        assert self.user_table.shape[1] == self.item_table.shape[1], "user and item vectors must have the same width"
        if self.bias is not None:
            self.bias = self.bias.reshape(-1)
            assert self.bias.shape[0] == self.item_table.shape[0]
        self.user_index = user_index
        self.note = note
        self.score_fn = "dot"  # "neg_l2": score = -||u - i|| + b (read by RankingEvaluator)
        self._torch = torch

    def _rows(self, users):
        torch = self._torch
        if self.user_index is not None:
            users = [self.user_index[u] for u in users]
        n = self.user_table.shape[0]
        if len(users) == n and (isinstance(users, range) or (users[0] == 0 and users[-1] == n - 1 and
                                                               np.array_equal(np.asarray(users), np.arange(n)))):
            return self.user_table
        idx = torch.as_tensor(np.asarray(users, dtype=np.int64), device=self.user_table.device)
        return self.user_table.index_select(0, idx)

    def eval_embeddings(self, users, item_shard=None):
        if item_shard is None:
            return self._rows(users), self.item_table, self.bias
        from . import dist
        rank, world = item_shard
        n_items = int(self.item_table.shape[0])
        lo, hi = dist.shard_range(n_items, rank, world)
        return self._rows(users), self.item_table[lo:hi], None if self.bias is None else self.bias[lo:hi], n_items

    def predict(self, users):
        if self.score_fn == "neg_l2":  # TransRec.py:90: -torch.norm(q.unsqueeze(1) - I, dim=-1)
            q = self._rows(users).float()
            s = -self._torch.norm(q.unsqueeze(1) - self.item_table.float().unsqueeze(0), p=None, dim=-1)
        else:
            s = self._rows(users).float() @ self.item_table.float().T
        if self.bias is not None:
            s = s + self.bias
        return s.cpu().numpy()

    def bind(self):
        """The bound `eval_embeddings`, to hang on an existing model object."""
        return self.eval_embeddings


def dot_product(user_table, item_table, item_bias=None, user_index=None):
    """score(u, i) = U[u] . I[i] (+ b[i])"""
    return EmbeddingScorer(user_table, item_table, item_bias, user_index, "dot")


def two_tower_sum(u_a, i_a, u_b, i_b, item_bias=None, user_index=None):
    """score = u_a.i_a + u_b.i_b  ==  [u_a | u_b] . [i_a | i_b]  (one contraction of width d_a + d_b)"""
    import torch
    u_a, i_a, u_b, i_b = _t(u_a), _t(i_a), _t(u_b), _t(i_b)
    return EmbeddingScorer(torch.cat([u_a, u_b], 1), torch.cat([i_a, i_b], 1), item_bias, user_index, "two_tower_sum")


def summed_query(parts, item_table, item_bias=None, user_index=None):
    """score = sum_p (q_p . i) + b_i  ==  (sum_p q_p) . i + b_i: several query vectors of one user against the same item
    table (HGN.py:147-163 adds user_emb.W2^T, union_out.W2^T and the sum over the sequence of item_embs.W2^T).
    The parts are added first, in the order given -- a float reassociation of the reference's sum of products."""
    parts = [_t(p).float() for p in parts]
    q = parts[0].clone()
    for p in parts[1:]:
        assert p.shape == q.shape, "query parts must have the same shape [B, d]"
        q += p
    return EmbeddingScorer(q, item_table, item_bias, user_index, "summed_query")


def decoder_layer(hidden_rows, weight, bias=None, user_index=None):
    """score = h_u . W[i, :] + c[i]: the last linear layer of an auto-encoder applied to the users' hidden
    rows (`hidden_rows[B, h]`, computed by the model for exactly the evaluated users)."""
    return EmbeddingScorer(hidden_rows, weight, bias, user_index, "decoder_layer")


def neg_euclidean(user_table, item_table, user_index=None):
    """score = -||u - i||.  Per user, -||u - i|| is a strictly decreasing function of
    ||u - i||^2 = ||u||^2 - 2 u.i + ||i||^2, so ranking by 2 u.i - ||i||^2 gives the same lists:
    user rows are doubled and -||i||^2 becomes the item bias.  (Scores differ; ranks and metrics do not,
    except where float rounding makes two distances equal.)"""
    u, i = _t(user_table).float(), _t(item_table).float()
    return EmbeddingScorer(2.0 * u, i, -(i * i).sum(1), user_index, "neg_euclidean(monotone)")


def neg_l2_plus_bias(query_table, item_table, item_bias=None, user_index=None):
    """score(u, i) = -||q_u - i|| + b_i (TransRec.py:86-93, SGAT.py:300).  `query_table[B, d]` holds the translated
    query of every evaluated user.  Scored by the FP32 tile kernels with a distance inner loop (see the module text)."""
    sc = EmbeddingScorer(query_table, item_table, item_bias, user_index, "neg_l2_plus_bias")
    sc.score_fn = "neg_l2"
    return sc


def transrec(user_table, global_transition, item_table, item_bias, last_items, user_index=None):
    """TransRec.py:86-93: q_u = U[u] + g + I[last_item(u)], score = -||q_u - i|| + b_i.  `last_items[u]` is the item the
    user interacted with last (`user_pos_dict[u][-1]`, TransRec.py:154), one per row of `user_table`."""
    import torch
    u, g, it = _t(user_table).float(), _t(global_transition).float().reshape(1, -1), _t(item_table).float()
    last = torch.as_tensor(np.asarray(last_items, dtype=np.int64), device=it.device)
    return neg_l2_plus_bias(u + g + it.index_select(0, last), it, item_bias, user_index)


def item_scores(scores, num_users, user_index=None):
    """The same score vector for every user (popularity baselines): a width-1 contraction 1 . s[i]."""
    import torch
    s = _t(scores).float().reshape(-1, 1)
    return EmbeddingScorer(torch.ones((int(num_users), 1), dtype=torch.float32), s, None, user_index, "item_scores")

"""TEST INFRASTRUCTURE - numpy restatement of the MNSRF ranking path (neuroir/multitask/mnsrf.py:61-162), float64,
plain loops: small cases only.  Pinned by tests/golden/mnsrf_*.npz (outputs of the unmodified reference,
oracle/gen_golden.py gen_session_ranker).  Only tests/ may import it."""
import numpy as np


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_dir(x, length, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of torch.nn.LSTM over one packed sequence (gate order i,f,g,o): x [L,in] -> bank [L,h], zeros at t >= length
    (rnn_encoder.py:73-76,110: pack / unpack); the reverse direction starts at the sequence's own last token."""
    L, h = x.shape[0], w_hh.shape[1]
    out = np.zeros((L, h))
    hs, cs = np.zeros(h), np.zeros(h)
    order = range(length - 1, -1, -1) if reverse else range(length)
    for t in order:
        g = w_ih @ x[t] + b_ih + w_hh @ hs + b_hh
        i, f, gg, o = _sig(g[:h]), _sig(g[h:2 * h]), np.tanh(g[2 * h:3 * h]), _sig(g[3 * h:])
        cs = f * cs + i * gg
        hs = o * np.tanh(cs)
        out[t] = hs
    return out


def encode(sd, prefix, emb, ids, lens):
    """(Bi)LSTM memory banks [n,L,H] of the sequences ids [n,L]."""
    p = prefix + '.encoder.rnns.0.'
    banks = []
    for r in range(ids.shape[0]):
        x = emb[ids[r]].astype(np.float64)
        parts = [lstm_dir(x, int(lens[r]), sd[p + 'weight_ih_l0'], sd[p + 'weight_hh_l0'], sd[p + 'bias_ih_l0'], sd[p + 'bias_hh_l0'], False)]
        if p + 'weight_ih_l0_reverse' in sd:
            parts.append(lstm_dir(x, int(lens[r]), sd[p + 'weight_ih_l0_reverse'], sd[p + 'weight_hh_l0_reverse'],
                                  sd[p + 'bias_ih_l0_reverse'], sd[p + 'bias_hh_l0_reverse'], True))
        banks.append(np.concatenate(parts, axis=1))
    return np.stack(banks)


def mnsrf_rank(sd, q, qlen, d, dlen):
    """q [B,S,Lq], qlen [B,S], d [B,S,N,Ld], dlen [B,S,N] -> (scores [B,S,N], memory_bank [B,S,Hq], session_bank [B,S,Hs])."""
    sd = {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}
    emb = sd['embedder.word_embeddings.make_embedding.emb_luts.0.weight']
    B, S, Lq = q.shape
    N, Ld = d.shape[2], d.shape[3]
    mem = encode(sd, 'query_encoder', emb, q.reshape(B * S, Lq), qlen.reshape(-1)).max(axis=1).reshape(B, S, -1)   # :75-80
    p = 'session_query_encoder.encoder.rnns.0.'
    sess = np.stack([lstm_dir(mem[b], S, sd[p + 'weight_ih_l0'], sd[p + 'weight_hh_l0'], sd[p + 'bias_ih_l0'], sd[p + 'bias_hh_l0'], False)
                     for b in range(B)])                                                                             # :84-105
    docs = encode(sd, 'document_encoder', emb, d.reshape(B * S * N, Ld), dlen.reshape(-1)).max(axis=1).reshape(B, S, N, -1)  # :127-131
    W, bias = sd['projection.linear.weight'], sd['projection.linear.bias']
    scores = np.zeros((B, S, N))
    for i in range(S):
        ctx = np.zeros_like(sess[:, 0]) if i == 0 else sess[:, i]                                                    # :143-146
        comb = np.tanh(np.concatenate([mem[:, i], ctx], axis=1) @ W.T + bias)                                        # :148
        scores[:, i] = np.einsum('bh,bnh->bn', comb, docs[:, i])                                                     # :155
    return scores, mem, sess

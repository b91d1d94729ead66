"""Training path of STAGE on the device: forward with saved activations, hand-written backward, clip + Adam.

The reference trains through PyTorch autograd (main.py:53-66: ``loss.backward()``, ``clip_grad_norm_``, ``Adam.step()``) and
``nn.DataParallel`` (main.py:204-206).  Here:

* ``TrainEngine`` runs the train-mode forward (dropout included, train branch of ``get_proposals``) out of the primitives of
  ``csrc/train.cuh`` and keeps what the backward needs; ``backward`` walks the same graph in reverse and accumulates every
  parameter gradient into ONE flat fp32 buffer (state_dict order) - the buffer a data-parallel run all-reduces.
* ``StageTrainFunction`` wraps the pair as a ``torch.autograd.Function`` so that the reference's own training loop
  (``outputs, att_loss, _, temporal_loss, _ = model(batch)`` ... ``loss.backward()``; ``optimizer.step()``) runs unchanged
  on ``tvqaplus_b200.model.STAGE`` in train mode.
* ``Trainer`` is the fused step for one-process-per-GPU training: forward, cls / temporal / attention losses and their
  gradients on the device (no host read), backward, one NCCL SUM all-reduce of the flat gradient buffer when
  ``torch.distributed`` is initialised, global-norm clip + Adam in one kernel on the flat parameter buffer.

Row order inside the engine: the attended tensors are (n, a, i, t) - question word before frame - so that every
contraction of the context-query attention (S = C Q^T, U = P Q and their four gradient products) and the reductions over
frames are plain strided batched GEMMs; the statement is permuted to the reference's (n, a, t, i) order once, in front of the
cls encoder.  All arithmetic is fp32 (CUDA-core FMA): gradients are checked against autograd of the reference formulas at
1e-4 (tests/test_train_gpu.py).
"""
from __future__ import annotations

import zlib
from typing import Dict

import torch

from . import ops
from . import train_ops as T


TC_MIN_ROWS = 4096  # below this the exact fp32 CUDA-core GEMM is as fast as a tensor-core launch


def _zeros(n, device):
    return torch.zeros(n, device=device, dtype=torch.float32)


def lin_fwd(x, w, b, relu=False):
    """y = x W^T + b: the split-fp16 tcgen05 kernel of the inference path (csrc/linear.cuh, fp32-grade: three fp16 products,
    fp32 accumulation) for the large row counts, the exact fp32 GEMM otherwise."""
    m, k = x.shape
    n = w.shape[0]
    if m >= TC_MIN_ROWS and k % 4 == 0 and k <= 768 and 8 <= n <= 304:
        return ops.linear_fwd(x, ops.pack_weight(w), b, n, relu=relu)
    return T.linear(x, w, b, relu=relu)


def lin_bwd_x(dy, w):
    """dx = dy W (W is [N, K]): the same tensor-core kernel with the transposed weight, column chunks of at most 256 outputs
    written as slices of dx.  Every gradient row is scaled by an exact power of two into fp16's normal range first
    (x_rowmax: one max pass over dy) - the rows are far below 6e-5 and orders of magnitude apart."""
    m, n = dy.shape
    k = w.numel() // n
    if not (m >= TC_MIN_ROWS and n % 4 == 0 and n <= 768 and k % 4 == 0):
        return T.linear_bwd_x(dy, w)
    wt = w.reshape(n, k).t().contiguous()  # [K, N]: row c = the weights of output column c of dx
    step = 128 if k % 128 == 0 and k <= 384 else (256 if k > 304 else k)
    dx = torch.empty(m, k, device=dy.device, dtype=torch.float32)
    rmax = T.row_absmax(dy)
    zero = _zeros(step, dy.device)
    for c0 in range(0, k, step):
        c1 = min(k, c0 + step)
        ops.linear_fwd(dy, ops.pack_weight(wt[c0:c1].contiguous()), zero[:c1 - c0], c1 - c0, out=dx, ldy=k, y_off=c0, x_rowmax=rmax)
    return dx


class TrainEngine:
    def __init__(self, model):
        self.model = model
        opt_ok = (model.sub_flag and model.vfeat_flag and model.add_local and model.t_iter == 0
                  and all(b.num_heads == 0 for e in (model.input_encoder, model.cls_encoder) for b in e.stacked_encoderBlocks))
        if not opt_ok:
            raise NotImplementedError("tvqaplus_b200: the training path is built for the run_main.sh configuration "
                                      "(sub + vfeat streams, add_local, t_iter = 0, encoder heads = 0)")
        self.names = [n for n, _ in model.named_parameters()]
        self.shapes = {n: tuple(p.shape) for n, p in model.named_parameters()}
        self.offsets, at = {}, 0
        for n, p in model.named_parameters():  # every tensor starts on a 256-byte boundary (vector loads, bulk copies)
            self.offsets[n] = at
            at += (p.numel() + 63) // 64 * 64
        self.n_params = at
        self.p_drop = float(model.dropout)
        self.seed_dev = None  # int32 [seed, forward-call counter] on the device (created at the first forward)

    # ---------------------------------------------------------------------------------------------- helpers
    def _drop(self, x, tag, out=None):
        """Dropout site ``tag`` (forward: on the activation; backward: the same call on its gradient - the mask is a pure
        function of (seed, call counter, tag, element index))."""
        if self.p_drop <= 0.0:
            return x
        return T.dropout(x, self.p_drop, self.seed_dev, zlib.crc32(tag.encode()), out=out)

    def _P(self) -> Dict[str, torch.Tensor]:
        return {n: p.detach() for n, p in self.model.named_parameters()}

    def new_grads(self, device):
        flat = torch.zeros(self.n_params, device=device, dtype=torch.float32)
        views = {n: flat[o:o + int(torch.Size(self.shapes[n]).numel())].view(self.shapes[n]) for n, o in self.offsets.items()}
        return flat, views

    # ---------------------------------------------------------------------------------------------- input projections
    def _fc_fwd(self, P, x, pre, tag):
        h0, s0 = T.ln_fwd(x, P[pre + ".0.weight"], P[pre + ".0.bias"])
        h0 = self._drop(h0, tag + ".fc")      # nn.Dropout of bert_word_encoding_fc / vid_fc (model/stage.py:87,100)
        r1 = lin_fwd(h0, P[pre + ".2.weight"], P[pre + ".2.bias"], relu=True)
        h1, s1 = T.ln_fwd(r1, P[pre + ".4.weight"], P[pre + ".4.bias"])
        h1 = self._drop(h1, tag + ".emb")     # nn.Dropout of input_embedding (model/stage.py:116)
        r2 = lin_fwd(h1, P["input_embedding.1.weight"], P["input_embedding.1.bias"], relu=True)
        e0, s2 = T.ln_fwd(r2, P["input_embedding.3.weight"], P["input_embedding.3.bias"])
        return e0, (x, h0, s0, r1, h1, s1, r2, s2, pre, tag)

    def _fc_bwd(self, P, G, de0, sv):
        x, h0, s0, r1, h1, s1, r2, s2, pre, tag = sv
        dz2 = T.ln_bwd(de0, r2, s2, P["input_embedding.3.weight"], G["input_embedding.3.weight"], G["input_embedding.3.bias"], relu_x=True)
        T.linear_bwd_w(dz2, h1, G["input_embedding.1.weight"], G["input_embedding.1.bias"])
        dh1 = self._drop(lin_bwd_x(dz2, P["input_embedding.1.weight"]), tag + ".emb")
        dz1 = T.ln_bwd(dh1, r1, s1, P[pre + ".4.weight"], G[pre + ".4.weight"], G[pre + ".4.bias"], relu_x=True)
        T.linear_bwd_w(dz1, h0, G[pre + ".2.weight"], G[pre + ".2.bias"])
        dh0 = self._drop(lin_bwd_x(dz1, P[pre + ".2.weight"]), tag + ".fc")
        T.ln_bwd(dh0, x, s0, P[pre + ".0.weight"], G[pre + ".0.weight"], G[pre + ".0.bias"], want_dx=False)

    # ---------------------------------------------------------------------------------------------- encoder blocks
    def _enc_fwd(self, P, x, n_seq, L, enc, prefix, tag):
        blocks = []
        cur_in = x
        ksize = enc.stacked_encoderBlocks[0].kernel_size
        for b, mod in enumerate(enc.stacked_encoderBlocks):
            blk = "%s.stacked_encoderBlocks.%d" % (prefix, b)
            pe = mod.position_encoding.pe[:L].contiguous()
            cur = T.add(cur_in, pe, period=L * 128)
            convs = []
            for i in range(mod.n_conv):
                nrm, st = T.ln_fwd(cur, P["%s.layer_norm.%d.weight" % (blk, i)], P["%s.layer_norm.%d.bias" % (blk, i)])
                if i % 2 == 0:  # F.dropout in front of every other conv (model/encoder.py:41-42)
                    nrm = self._drop(nrm, "%s.enc.%d.%d" % (tag, b, i))
                dw = T.dwconv_fwd(nrm, P["%s.conv.%d.depthwise_conv.weight" % (blk, i)], P["%s.conv.%d.depthwise_conv.bias" % (blk, i)],
                                  n_seq, L, ksize)
                pw = lin_fwd(dw, P["%s.conv.%d.pointwise_conv.weight" % (blk, i)].reshape(128, 128), P["%s.conv.%d.pointwise_conv.bias" % (blk, i)], relu=True)
                nxt = T.add(pw, cur)
                convs.append((cur, st, nrm, dw, pw))
                cur = nxt
            out, stf = T.ln_fwd(cur, P[blk + ".final_layer_norm.weight"], P[blk + ".final_layer_norm.bias"])
            blocks.append((convs, cur, stf, blk, b))
            cur_in = out
        return cur_in, (blocks, n_seq, L, ksize, tag)

    def _enc_bwd(self, P, G, dout, sv):
        blocks, n_seq, L, k, tag = sv
        d = dout
        for convs, last, stf, blk, b in reversed(blocks):
            d = T.ln_bwd(d, last, stf, P[blk + ".final_layer_norm.weight"], G[blk + ".final_layer_norm.weight"], G[blk + ".final_layer_norm.bias"])
            for i in reversed(range(len(convs))):
                cur, st, nrm, dw, pw = convs[i]
                pwn, dwn, lnn = "%s.conv.%d.pointwise_conv" % (blk, i), "%s.conv.%d.depthwise_conv" % (blk, i), "%s.layer_norm.%d" % (blk, i)
                dpw = T.relu_bwd(d, pw)
                T.linear_bwd_w(dpw, dw, G[pwn + ".weight"], G[pwn + ".bias"])
                ddw = lin_bwd_x(dpw, P[pwn + ".weight"].reshape(128, 128))
                dn = T.dwconv_bwd(ddw, nrm, P[dwn + ".weight"], G[dwn + ".weight"], G[dwn + ".bias"], n_seq, L, k)
                if i % 2 == 0:
                    dn = self._drop(dn, "%s.enc.%d.%d" % (tag, b, i))
                d = T.ln_bwd(dn, cur, st, P[lnn + ".weight"], G[lnn + ".weight"], G[lnn + ".bias"], add=d, out=d)
        return d  # the position table has no gradient

    # ---------------------------------------------------------------------------------------------- context-query attention
    def _cqa_fwd(self, P, C, Cn, c_mask, Q, q_mask, N, Tn, Lq, Lc, scale, tag):
        ai, dev = 5 * Lq, C.device
        Qn, qinv = T.l2norm_fwd(Q)
        # F.dropout on the NORMALISED operands of the similarity only (context_query_attention.py:95-96); U = P Q uses Q itself
        Cd = self._drop(Cn, tag + ".cn", out=torch.empty_like(Cn)) if self.p_drop > 0 else Cn
        Qd = self._drop(Qn, tag + ".qn", out=torch.empty_like(Qn)) if self.p_drop > 0 else Qn
        S = torch.empty(N * ai * Tn, Lc, device=dev, dtype=torch.float32)
        T.gemm(Cd, Qd, S, ai, Tn * Lc, 128, (128, 1), (1, 128), Tn * Lc, batch=(N, 1), a_b=(ai * 128, 0), b_b=(Tn * Lc * 128, 0),
               c_b=(ai * Tn * Lc, 0))
        Pm = T.attn_softmax_fwd(S, c_mask, q_mask, N, ai, Tn, Lc, scale)
        U = torch.empty(N * ai * Tn, 128, device=dev, dtype=torch.float32)
        T.gemm(Pm, Q, U, ai, 128, Lc, (Tn * Lc, 1), (128, 1), Tn * 128, batch=(N, Tn), a_b=(ai * Tn * Lc, Lc), b_b=(Tn * Lc * 128, Lc * 128),
               c_b=(ai * Tn * 128, 128))
        cat = T.cat3_fwd(C, U, Tn)
        l, st = T.ln_fwd(cat, P["c2q_down_projection.0.weight"], P["c2q_down_projection.0.bias"])
        l = self._drop(l, tag + ".c2q")  # nn.Dropout of c2q_down_projection (model/stage.py:135)
        att = lin_fwd(l, P["c2q_down_projection.2.weight"], P["c2q_down_projection.2.bias"], relu=True)
        return att, S, (C, Cn, Cd, Q, Qn, Qd, qinv, S, Pm, U, cat, l, st, att, N, Tn, Lq, Lc, scale, tag)

    def _cqa_bwd(self, P, G, datt, sv, dC, cinv, att_grad=None):
        C, Cn, Cd, Q, Qn, Qd, qinv, S, Pm, U, cat, l, st, att, N, Tn, Lq, Lc, scale, tag = sv
        ai, dev = 5 * Lq, C.device
        dpre = T.relu_bwd(datt, att, out=datt)
        T.linear_bwd_w(dpre, l, G["c2q_down_projection.2.weight"], G["c2q_down_projection.2.bias"])
        dl = self._drop(lin_bwd_x(dpre, P["c2q_down_projection.2.weight"]), tag + ".c2q")
        dcat = T.ln_bwd(dl, cat, st, P["c2q_down_projection.0.weight"], G["c2q_down_projection.0.weight"], G["c2q_down_projection.0.bias"], out=dl)
        dU = T.cat3_bwd(dcat, C, U, dC, Tn, True)
        del dcat, dl
        dP = torch.empty_like(Pm)
        T.gemm(dU, Q, dP, ai, Lc, 128, (Tn * 128, 1), (1, 128), Tn * Lc, batch=(N, Tn), a_b=(ai * Tn * 128, 128), b_b=(Tn * Lc * 128, Lc * 128),
               c_b=(ai * Tn * Lc, Lc))
        dQ = torch.empty_like(Q)
        T.gemm(Pm, dU, dQ, Lc, 128, ai, (1, Tn * Lc), (Tn * 128, 1), 128, batch=(N, Tn), a_b=(ai * Tn * Lc, Lc), b_b=(ai * Tn * 128, 128),
               c_b=(Tn * Lc * 128, Lc * 128))
        dS = T.attn_softmax_bwd_(dP, Pm, Lc, scale)
        if att_grad is not None:  # the ranking loss reads the masked raw scores of the vid stream (model/stage.py:317-341)
            att_grad(S, dS)
        dCn = torch.zeros_like(C)
        T.gemm(dS, Qd, dCn, ai, 128, Tn * Lc, (Tn * Lc, 1), (128, 1), 128, batch=(N, 1), a_b=(ai * Tn * Lc, 0), b_b=(Tn * Lc * 128, 0),
               c_b=(ai * 128, 0), mode=2, split_k=max(1, min(64, 592 // max(1, N * ((ai + 127) // 128)))))
        T.l2norm_bwd(self._drop(dCn, tag + ".cn"), Cn, cinv, dC, True)
        dQn = torch.empty_like(Q)
        T.gemm(dS, Cd, dQn, Tn * Lc, 128, ai, (1, Tn * Lc), (128, 1), 128, batch=(N, 1), a_b=(ai * Tn * Lc, 0), b_b=(ai * 128, 0),
               c_b=(Tn * Lc * 128, 0))
        T.l2norm_bwd(self._drop(dQn, tag + ".qn"), Qn, qinv, dQ, True)
        return dQ

    # ---------------------------------------------------------------------------------------------- forward
    def forward(self, batch, att_offsets=None) -> dict:
        """batch: the reference's model inputs on the device.  att_offsets = (pos, neg) int64 element offsets of the sampled
        attention pairs in the ENGINE's score layout (see ``att_offsets_from_pairs``), or None (no attention loss).
        Returns the outputs and ``ctx`` for ``backward``."""
        m = self.model
        P = self._P()
        N, Lq = batch.qas_bert.shape[0], batch.qas_bert.shape[2]
        Tn, Ls = batch.sub_bert.shape[1:3]
        R = batch.vid.shape[2]
        qas_mask = batch.qas_mask.float().contiguous()
        sub_mask = batch.sub_mask.float().contiguous()
        vid_mask = batch.vid_mask.float().contiguous()
        ts_mask = batch.ts_label_mask.float().contiguous()
        target = batch.target.to(torch.int64).contiguous()
        st_lab, ed_lab = batch.ts_label["st"].to(torch.int64).contiguous(), batch.ts_label["ed"].to(torch.int64).contiguous()
        ctx = {}
        if self.p_drop > 0:  # a new mask set per forward call: advance the call counter next to the seed
            if self.seed_dev is None:
                d = torch.distributed
                rank = d.get_rank() if d.is_available() and d.is_initialized() else 0  # data-parallel ranks draw different masks
                self.seed_dev = torch.tensor([(int(torch.initial_seed()) + 7919 * rank) & 0x7FFFFFFF, 0], dtype=torch.int32,
                                             device=batch.qas_bert.device)
            T.counter_inc(self.seed_dev[1:])

        # 1. input projections + input encoder, three streams (model/stage.py:351-363)
        qa0, ctx["fc_q"] = self._fc_fwd(P, batch.qas_bert.float().reshape(N * 5 * Lq, -1), "bert_word_encoding_fc", "q")
        C, ctx["enc_q"] = self._enc_fwd(P, qa0, N * 5, Lq, m.input_encoder, "input_encoder", "q")
        sub0, ctx["fc_s"] = self._fc_fwd(P, batch.sub_bert.float().reshape(N * Tn * Ls, -1), "bert_word_encoding_fc", "s")
        Qs, ctx["enc_s"] = self._enc_fwd(P, sub0, N * Tn, Ls, m.input_encoder, "input_encoder", "s")
        vn, _ = T.l2norm_fwd(batch.vid.float().reshape(N * Tn * R, -1), want_inv=False)  # model/stage.py:256
        vid0, ctx["fc_v"] = self._fc_fwd(P, vn, "vid_fc", "v")
        Qv, ctx["enc_v"] = self._enc_fwd(P, vid0, N * Tn, R, m.input_encoder, "input_encoder", "v")
        del qa0, sub0, vid0, vn

        # 2. context-query attention + c2q projection per stream, rows (n, a, i, t) (model/stage.py:365-387)
        Cn, cinv = T.l2norm_fwd(C)
        ctx["c"] = (C, Cn, cinv)
        att_s, _, ctx["cqa_s"] = self._cqa_fwd(P, C, Cn, qas_mask, Qs, sub_mask, N, Tn, Lq, Ls, m.scale, "s")
        att_v, S_vid, ctx["cqa_v"] = self._cqa_fwd(P, C, Cn, qas_mask, Qv, vid_mask, N, Tn, Lq, R, m.scale, "v")

        # 3. concat_fc (model/stage.py:276-279), then the reference's (n, a, t, i) row order for the cls encoder
        cat2 = T.cat3_fwd(att_s, att_v, 1)
        l2, st2 = T.ln_fwd(cat2, P["concat_fc.0.weight"], P["concat_fc.0.bias"])
        l2 = self._drop(l2, "cat")  # nn.Dropout of concat_fc (model/stage.py:109)
        r = lin_fwd(l2, P["concat_fc.2.weight"], P["concat_fc.2.bias"], relu=True)
        stm, st4 = T.ln_fwd(r, P["concat_fc.4.weight"], P["concat_fc.4.bias"])
        ctx["cat"] = (att_s, att_v, cat2, l2, st2, r, st4)
        stm_p = T.permute_it(stm, N * 5, Lq, Tn)
        del stm
        frame_on = (vid_mask.view(N, Tn, R).sum(-1) != 0).float()                      # statement mask = the vid stream's (stage.py:280-282)
        sm = (qas_mask.view(N, 5, 1, Lq) * frame_on.view(N, 1, Tn, 1)).contiguous()     # (N, 5, T, Lq)
        mm = (sm.sum(-1) != 0).float().contiguous()                                     # max_statement_mask (N, 5, T)

        # 4. cls encoder + masked max over the question words (model/stage.py:499-505)
        enc, ctx["enc_c"] = self._enc_fwd(P, stm_p, N * 5 * Tn, Lq, m.cls_encoder, "cls_encoder", "c")
        ms, arg = T.maxpool_fwd(enc, sm.view(-1, Lq), N * 5 * Tn, Lq)
        del enc, stm_p

        # 5. residual temporal predictor, span scores, temporal loss (model/stage.py:469-482,521,539-555)
        pj, stl, edl, cl = "cls_projection_layers.0.conv", "temporal_scoring_st_layers.0.conv", "temporal_scoring_ed_layers.0.conv", "classifier.conv"
        l0, s0 = T.ln_fwd(ms, P[pj + ".0.weight"], P[pj + ".0.bias"])
        l0 = self._drop(l0, "proj")  # LinearWrapper: LayerNorm -> Dropout -> Linear (model/stage.py:15-32)
        p0 = lin_fwd(l0, P[pj + ".2.weight"], P[pj + ".2.bias"], relu=True)
        x1 = T.add(ms, p0)
        ls, sst = T.ln_fwd(x1, P[stl + ".0.weight"], P[stl + ".0.bias"])
        le, sed = T.ln_fwd(x1, P[edl + ".0.weight"], P[edl + ".0.bias"])
        ls, le = self._drop(ls, "st"), self._drop(le, "ed")
        ts = torch.empty(N * 5 * Tn, 2, device=ms.device, dtype=torch.float32)
        T.linear(ls, P[stl + ".2.weight"], P[stl + ".2.bias"], out=ts, ldc=2, c_off=0)
        T.linear(le, P[edl + ".2.weight"], P[edl + ".2.bias"], out=ts, ldc=2, c_off=1)
        t_scores = ops.mask_logits_fwd(ts.view(N, 5, Tn, 2), ts_mask.view(N, 1, Tn).expand(N, 5, Tn).contiguous())
        terms, probs, dts = T.ts_loss_grad(t_scores, ts_mask, target, st_lab, ed_lab)
        temporal_loss = terms.sum() * 0.5

        # 6. train branch of get_proposals + classifier (model/stage.py:407-438,536)
        spans, conf = ops.find_max_span(probs[0], probs[1])
        feat, parg, valid, row_target = T.proposal_pool_fwd(x1, mm, target, st_lab, ed_lab, spans, conf, N, Tn, m.extra_span_length, 0.01)
        lc, sc = T.ln_fwd(feat, P[cl + ".0.weight"], P[cl + ".0.bias"])
        lc = self._drop(lc, "cls")
        rows = T.linear(lc, P[cl + ".2.weight"], P[cl + ".2.bias"]).view(2 * N, 5)
        self.last_row_valid = valid  # (2N,) int32: which proposal rows exist (row 2n = GT span, 2n+1 = predicted span)
        ctx["head"] = (ms, arg, sm, mm, l0, s0, p0, x1, ls, sst, le, sed, dts, feat, parg, valid, lc, sc, N, Tn, Lq)

        # 7. attention ranking loss on the vid stream's raw scores (model/stage.py:317-341,731-746)
        att_loss = None
        if att_offsets is not None:
            pos, neg = (o.to(device=S_vid.device, dtype=torch.int64).contiguous() for o in att_offsets)
            att_loss = ops.att_loss_fwd(S_vid, pos, neg, loss_type=m.att_loss_type, alpha=m.alpha, margin=m.margin)
            ctx["att"] = (pos, neg)
        return dict(answer_rows=rows, row_valid=valid, row_target=row_target, t_scores=t_scores, temporal_loss=temporal_loss,
                    att_loss=att_loss, vid_raw_s=S_vid, ctx=ctx)

    def att_offsets_from_pairs(self, pos, neg, N, Tn, Lq, R):
        """(batch, answer, image, word, region) index rows of ``STAGE._att_pairs`` -> element offsets into the engine's
        (N, 5, Lq, T, R) score tensor."""
        strides = torch.tensor([5 * Lq * Tn * R, Lq * Tn * R, R, Tn * R, 1], dtype=torch.long)
        return (pos * strides).sum(1), (neg * strides).sum(1)

    # ---------------------------------------------------------------------------------------------- backward
    def backward(self, ctx, d_rows, g_ts, g_att=None):
        """d_rows (2N, 5): gradient w.r.t. ``answer_rows`` (zero on invalid rows); g_ts / g_att: one-element device tensors,
        the upstream gradients of ``temporal_loss`` / ``att_loss``.  Returns the flat gradient buffer (state_dict order)."""
        m = self.model
        P = self._P()
        dev = d_rows.device
        flat, G = self.new_grads(dev)
        ms, arg, sm, mm, l0, s0, p0, x1, ls, sst, le, sed, dts, feat, parg, valid, lc, sc, N, Tn, Lq = ctx["head"]
        pj, stl, edl, cl = "cls_projection_layers.0.conv", "temporal_scoring_st_layers.0.conv", "temporal_scoring_ed_layers.0.conv", "classifier.conv"

        # classifier and the pooled proposal features
        dr = d_rows.contiguous().view(-1, 1)
        T.linear_bwd_w(dr, lc, G[cl + ".2.weight"], G[cl + ".2.bias"])
        dlc = self._drop(T.linear_bwd_x(dr, P[cl + ".2.weight"]), "cls")
        dfeat = T.ln_bwd(dlc, feat, sc, P[cl + ".0.weight"], G[cl + ".0.weight"], G[cl + ".0.bias"])
        dx1 = torch.zeros_like(x1)
        T.proposal_pool_bwd(dfeat, parg, valid, mm, N, Tn, dx1)
        # span scores (temporal loss)
        dts = T.scale_(dts.view(-1, 2).clone(), g_ts)
        T.linear_bwd_w(dts, ls, G[stl + ".2.weight"], G[stl + ".2.bias"], dy_ld=2, dy_off=0, n=1)
        dls = self._drop(T.linear_bwd_x(dts, P[stl + ".2.weight"], dy_ld=2, dy_off=0, n=1), "st")
        T.ln_bwd(dls, x1, sst, P[stl + ".0.weight"], G[stl + ".0.weight"], G[stl + ".0.bias"], add=dx1, out=dx1)
        T.linear_bwd_w(dts, le, G[edl + ".2.weight"], G[edl + ".2.bias"], dy_ld=2, dy_off=1, n=1)
        dle = self._drop(T.linear_bwd_x(dts, P[edl + ".2.weight"], dy_ld=2, dy_off=1, n=1), "ed")
        T.ln_bwd(dle, x1, sed, P[edl + ".0.weight"], G[edl + ".0.weight"], G[edl + ".0.bias"], add=dx1, out=dx1)
        # residual projection
        dp0 = T.relu_bwd(dx1, p0)
        T.linear_bwd_w(dp0, l0, G[pj + ".2.weight"], G[pj + ".2.bias"])
        dl0 = self._drop(lin_bwd_x(dp0, P[pj + ".2.weight"]), "proj")
        dms = T.ln_bwd(dl0, ms, s0, P[pj + ".0.weight"], G[pj + ".0.weight"], G[pj + ".0.bias"], add=dx1)

        # max-pool, cls encoder, back to (n, a, i, t)
        denc = T.maxpool_bwd(dms, sm.view(-1, Lq), arg, N * 5 * Tn, Lq)
        dstm_p = self._enc_bwd(P, G, denc, ctx["enc_c"])
        dstm = T.permute_it(dstm_p, N * 5, Tn, Lq)
        del dstm_p, denc

        # concat_fc
        att_s, att_v, cat2, l2, st2, r, st4 = ctx["cat"]
        dr2 = T.ln_bwd(dstm, r, st4, P["concat_fc.4.weight"], G["concat_fc.4.weight"], G["concat_fc.4.bias"], relu_x=True, out=dstm)
        T.linear_bwd_w(dr2, l2, G["concat_fc.2.weight"], G["concat_fc.2.bias"])
        dl2 = self._drop(lin_bwd_x(dr2, P["concat_fc.2.weight"]), "cat")
        dcat2 = T.ln_bwd(dl2, cat2, st2, P["concat_fc.0.weight"], G["concat_fc.0.weight"], G["concat_fc.0.bias"], out=dl2)
        datt_s = torch.empty_like(att_s)
        datt_v = T.cat3_bwd(dcat2, att_s, att_v, datt_s, 1, False)
        del dcat2, dl2, dr2, dstm

        # context-query attention, both streams; dC / dCn collect both
        C, Cn, cinv = ctx["c"]
        dC = torch.zeros_like(C)
        att_grad = None
        if g_att is not None and "att" in ctx:
            pos, neg = ctx["att"]

            def att_grad(S, dS):
                T.att_loss_bwd(S, pos, neg, 0 if m.att_loss_type == "lse" else 1, m.alpha, m.margin, g_att, dS)
        dQv = self._cqa_bwd(P, G, datt_v, ctx["cqa_v"], dC, cinv, att_grad)
        dQs = self._cqa_bwd(P, G, datt_s, ctx["cqa_s"], dC, cinv)

        # input encoders and projections
        for dq, e, f in ((dQv, "enc_v", "fc_v"), (dQs, "enc_s", "fc_s"), (dC, "enc_q", "fc_q")):
            self._fc_bwd(P, G, self._enc_bwd(P, G, dq, ctx[e]), ctx[f])
        return flat


class StageTrainFunction(torch.autograd.Function):
    """``(answer_rows[valid], att_loss, temporal_loss, t_scores) = StageTrainFunction.apply(engine, batch, att_offsets, *params)``
    with a backward that runs the engine's kernels and hands every parameter its gradient."""

    @staticmethod
    def forward(actx, engine, batch, att_offsets, *params):
        out = engine.forward(batch, att_offsets)
        keep = out["row_valid"].bool()  # N_new is data dependent: the one host sync of the train-mode forward
        idx = torch.nonzero(keep).squeeze(1)
        actx.engine, actx.saved, actx.idx, actx.n_rows = engine, out["ctx"], idx, keep.numel()
        actx.has_att = out["att_loss"] is not None
        att = out["att_loss"] if actx.has_att else torch.zeros((), device=keep.device)
        actx.mark_non_differentiable(out["t_scores"])
        return out["answer_rows"][idx], att, out["temporal_loss"], out["t_scores"], out["row_target"][idx]

    @staticmethod
    def backward(actx, d_rows, d_att, d_ts, _d_scores, _d_target):
        eng = actx.engine
        if actx.saved is None:
            raise RuntimeError("tvqaplus_b200: backward through the STAGE training graph a second time (the saved activations "
                               "are released after the first backward; run the forward again)")
        dev = actx.idx.device
        full = torch.zeros(actx.n_rows, 5, device=dev, dtype=torch.float32)
        if d_rows is not None:
            full[actx.idx] = d_rows.float()
        g_ts = (d_ts if d_ts is not None else torch.zeros((), device=dev)).reshape(1).float().contiguous()
        g_att = d_att.reshape(1).float().contiguous() if (actx.has_att and d_att is not None) else None
        flat = eng.backward(actx.saved, full, g_ts, g_att)
        actx.saved = None
        grads = tuple(flat[o:o + int(torch.Size(eng.shapes[n]).numel())].view(eng.shapes[n]) for n, o in eng.offsets.items())
        return (None, None, None) + grads


class Trainer:
    """One-process-per-GPU training step: ``step(batch)`` = forward, losses, backward, gradient all-reduce (when
    ``torch.distributed`` is initialised), global-norm clip + Adam (main.py:53-66, 204-212) without a host read.

    The parameters are re-pointed at views of one flat buffer (like DDP's buckets) so that clip + Adam is one kernel and the
    gradient of a step is one contiguous tensor - the single collective of data-parallel training (SUM: the losses are sums
    over examples, so the summed gradient is the large-batch gradient)."""

    def __init__(self, model, lr: float = 1e-3, wd: float = 3e-7, clip: float = 10.0, att_weight: float = 0.1, ts_weight: float = 0.5,
                 betas=(0.9, 0.999), eps: float = 1e-8):
        self.model, self.engine = model, TrainEngine(model)
        dev = next(model.parameters()).device
        self.flat = torch.zeros(self.engine.n_params, device=dev, dtype=torch.float32)
        for n, p in model.named_parameters():
            o = self.engine.offsets[n]
            self.flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.flat[o:o + p.numel()].view(p.shape)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.lr, self.wd, self.clip, self.betas, self.eps = lr, wd, clip, betas, eps
        self.g_ts = torch.full((1,), ts_weight, device=dev)
        self.g_att = torch.full((1,), att_weight, device=dev)
        self.att_weight, self.ts_weight = att_weight, ts_weight
        self.steps = 0
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)  # the step count Adam's bias corrections use
        self.last_grad = None
        self._graphs = {}

    # ---- one step = part A (forward, losses, backward -> the flat gradient) + [all-reduce] + part B (clip + Adam)
    @staticmethod
    def _world() -> int:
        d = torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    def _part_a(self, batch, att_offsets):
        eng = self.engine
        out = eng.forward(batch, att_offsets)
        n = batch.qas_bert.shape[0]
        cls, d_rows = T.answer_ce(out["answer_rows"], out["row_valid"], out["row_target"], n)
        grad = eng.backward(out["ctx"], d_rows, self.g_ts, self.g_att if out["att_loss"] is not None else None)
        loss = cls[0] + self.ts_weight * out["temporal_loss"]
        if out["att_loss"] is not None:
            loss = loss + self.att_weight * out["att_loss"]
        return dict(loss=loss, cls_loss=cls[0], n_rows=cls[1], n_correct=cls[2], temporal_loss=out["temporal_loss"],
                    att_loss=out["att_loss"], grad=grad)

    def _part_b(self, out):
        grad = out["grad"]
        ss = torch.zeros(1, device=grad.device)
        T.grad_sumsq(grad, ss)
        T.counter_inc(self.step_dev)
        T.adam_step(self.flat, grad, self.m, self.v, ss, max_norm=self.clip, lr=self.lr, beta1=self.betas[0], beta2=self.betas[1],
                    eps=self.eps, wd=self.wd, step_dev=self.step_dev)
        out["grad_sumsq"] = ss
        return out

    def _all_reduce(self, grad):
        if self._world() > 1:  # SUM: the three losses are sums over examples, so this is the large-batch gradient
            torch.distributed.all_reduce(grad, op=torch.distributed.ReduceOp.SUM)

    def _step_core(self, batch, att_offsets):
        out = self._part_a(batch, att_offsets)
        self._all_reduce(out["grad"])
        return self._part_b(out)

    _STATIC = ("qas_bert", "sub_bert", "vid", "qas_mask", "sub_mask", "vid_mask", "ts_label_mask", "target")

    def _capture(self, batch, att_offsets):
        """Record the step into CUDA graphs over static copies of the inputs; nothing executes during the capture.  One process:
        ONE graph (forward, losses, backward, clip + Adam).  Data-parallel: TWO graphs with the NCCL all-reduce launched
        eagerly between their replays - the collective stays out of the capture (the one attempt with the all-reduce inside
        the captured graph, 8 ranks, deadlocked)."""
        from .host import Batch
        static = Batch(batch)
        for k in self._STATIC:
            static[k] = batch[k].clone()
        static["ts_label"] = {k: v.clone() for k, v in batch["ts_label"].items()}
        offs = None if att_offsets is None else tuple(o.to(self.flat.device, torch.int64).contiguous().clone() for o in att_offsets)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        if self._world() == 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._step_core(static, offs)
            return static, offs, (g,), out
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga, capture_error_mode="thread_local"):  # the process group's watchdog thread keeps running
            out = self._part_a(static, offs)
        with torch.cuda.graph(gb, pool=ga.pool(), capture_error_mode="thread_local"):
            out = self._part_b(out)
        return static, offs, (ga, gb), out

    def step(self, batch, att_offsets=None, graph: bool = False) -> dict:
        """``graph=True``: the step is replayed from a CUDA graph captured for this batch geometry (the first call with a new
        geometry runs eagerly and records the graph for the following ones).  The returned tensors of a replay are the
        graph's static outputs: read them before the next step."""
        if not graph:
            out = self._step_core(batch, att_offsets)
        else:
            key = tuple(tuple(batch[k].shape) for k in self._STATIC) + (None if att_offsets is None else int(att_offsets[0].numel()),)
            rec = self._graphs.get(key)
            if rec is None:
                out = self._step_core(batch, att_offsets)
                self._graphs[key] = self._capture(batch, att_offsets)
            else:
                static, offs, graphs, out = rec
                for k in self._STATIC:
                    if static[k].data_ptr() != batch[k].data_ptr():
                        static[k].copy_(batch[k], non_blocking=True)
                for k, v in batch["ts_label"].items():
                    static["ts_label"][k].copy_(v, non_blocking=True)
                if offs is not None:
                    for dst, src in zip(offs, att_offsets):
                        dst.copy_(src, non_blocking=True)
                graphs[0].replay()
                if len(graphs) == 2:
                    self._all_reduce(out["grad"])
                    graphs[1].replay()
        self.steps += 1
        self.model.invalidate_weight_cache()  # the eval kernels' packed weight images are keyed on the version counter
        self.last_grad = out["grad"]
        return out

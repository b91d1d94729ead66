"""STAGE (reference model/stage.py:55-806) on hand-written sm_100a kernels.

Same constructor (``STAGE(opt)``), ``forward(batch)`` contract, sub-module names and ``state_dict`` keys as the reference,
so ``main.py`` / ``inference.py`` can import this class instead of ``model.stage.STAGE``.  The eval / inference forward is
sixteen kernel launches (csrc/*.cuh):

    linear x2 (text)  ->  encoder        a_embed   (5N, Lq, 128)
    linear x2 (text)  ->  encoder        sub_embed (N*T, Ls, 128)        [written straight into the CQA operand images]
    linear x2 (vid)   ->  encoder        vid_embed (N*T, R, 128)
    cqa_prep, a.Wa, cqa_fused            statement (N, 5, T, Lq, 128)   [S, P, U, both mixed tensors never leave the SM]
    fill, encoder (cls, masked max-pool) max_statement (5N*T, 128)
    span_head, ts_loss_sum               t_scores, spans, answer, temporal loss        [no device->host sync]

Train mode with gradients enabled runs the training engine instead (``tvqaplus_b200/train.py``: forward with saved
activations, dropout, train branch of ``get_proposals``; the returned tensors carry an autograd graph whose backward is the
hand-written backward kernels, so the reference's ``loss.backward()`` / ``optimizer.step()`` loop works unchanged).  Train
mode under ``torch.no_grad()`` runs the fused forward kernels plus the train branch of ``get_proposals`` (dropout must be 0).
What is not built raises ``NotImplementedError`` (never a silent CPU path): ``add_non_visual``, hard negatives, and - in
the training engine - attention heads, ``t_iter > 0``, single-stream models.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from . import _packing
from ._packing import packed, packed_ln
from .cnn import DepthwiseSeparableConv
from .context_query_attention import StructuredAttention
from .encoder import StackedEncoder, _no_train_dropout
from .model_utils import find_max_triples  # noqa: F401  (re-exported like the reference module does)


def _ln_pair(ln: nn.LayerNorm):
    return (ln.weight, ln.bias)


class LinearWrapper(nn.Module):
    """[LayerNorm] -> Dropout -> Linear [-> ReLU] over the last axis (reference model/stage.py:15-32); the Sequential is
    called ``conv`` there and here, so keys are ``conv.0.*`` (LayerNorm) and ``conv.2.*`` (Linear)."""

    def __init__(self, in_hsz, out_hsz, layer_norm=True, dropout=0.1, relu=True):
        super().__init__()
        self.relu = relu
        self.p = dropout
        mods = ([nn.LayerNorm(in_hsz)] if layer_norm else []) + [nn.Dropout(dropout), nn.Linear(in_hsz, out_hsz)]
        self.conv = nn.Sequential(*mods)

    def forward(self, x):
        _no_train_dropout(self, self.p)
        lin = self.conv[-1]
        ln = self.conv[0] if isinstance(self.conv[0], nn.LayerNorm) else None
        kc = ops.kc_for(lin.out_features)
        if ln is None:
            return ops.linear_fwd(x, packed(lin.weight, kc), lin.bias, lin.out_features, relu=self.relu)
        img, cs, cb = packed_ln(lin, ln, kc)
        return ops.linear_fwd(x, img, None, lin.out_features, ln_in=(cs, cb), relu=self.relu)


class ConvLinear(nn.Module):
    """[LayerNorm] -> Dropout -> DepthwiseSeparableConv (reference model/stage.py:35-52)."""

    def __init__(self, in_hsz, out_hsz, kernel_size=3, layer_norm=True, dropout=0.1, relu=True):
        super().__init__()
        self.p = dropout
        self.kernel_size = kernel_size
        mods = ([nn.LayerNorm(in_hsz)] if layer_norm else []) + [
            nn.Dropout(dropout), DepthwiseSeparableConv(in_ch=in_hsz, out_ch=out_hsz, k=kernel_size, dim=1, relu=relu)]
        self.conv = nn.Sequential(*mods)

    def forward(self, x):
        _no_train_dropout(self, self.p)
        dws = self.conv[-1]
        ln = self.conv[0] if isinstance(self.conv[0], nn.LayerNorm) else None
        return ops.encoder_block_fwd(x, None, None, [dws.conv_args(ln)], None, self.kernel_size,
                                     conv_only=ln is None, pre_ln_only=ln is not None, relu=dws.relu)


class STAGE(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.inference_mode = False
        self._last_ts_loss = None
        self.error_checks = True  # watch the device error flag (no synchronisation; see _forward_eval)
        self.sub_flag = opt.sub_flag
        self.vfeat_flag = opt.vfeat_flag
        self.vfeat_size = opt.vfeat_size
        self.t_iter = opt.t_iter
        self.extra_span_length = opt.extra_span_length
        self.add_local = opt.add_local
        self.use_sup_att = opt.use_sup_att
        self.num_negatives = opt.num_negatives
        self.negative_pool_size = opt.negative_pool_size
        self.num_hard = opt.num_hard
        self.drop_topk = opt.drop_topk
        self.margin = opt.margin
        self.att_loss_type = opt.att_loss_type
        self.scale = opt.scale
        self.alpha = opt.alpha
        self.dropout = opt.dropout
        self.hsz = opt.hsz
        self.bsz = None
        self.num_seg = None
        self.num_a = 5
        self.flag_cnt = self.sub_flag + self.vfeat_flag
        self.wd_size = opt.embedding_size
        self.bridge_hsz = 300
        if self.hsz != ops.HSZ:
            raise NotImplementedError("tvqaplus_b200: kernels are specialised for hsz=128 (config.py:103 default)")
        if getattr(opt, "add_non_visual", False):
            raise NotImplementedError("tvqaplus_b200: add_non_visual is not built (no void vectors exist in the reference either)")

        p = self.dropout
        # Module creation order follows the reference (model/stage.py:85-186) so that default initialisation under the
        # same torch seed yields the same weights.
        self.bert_word_encoding_fc = nn.Sequential(nn.LayerNorm(self.wd_size), nn.Dropout(p),
                                                   nn.Linear(self.wd_size, self.bridge_hsz), nn.ReLU(True),
                                                   nn.LayerNorm(self.bridge_hsz))
        if self.vfeat_flag:
            self.vid_fc = nn.Sequential(nn.LayerNorm(self.vfeat_size), nn.Dropout(p),
                                        nn.Linear(self.vfeat_size, self.bridge_hsz), nn.ReLU(True),
                                        nn.LayerNorm(self.bridge_hsz))
        if self.flag_cnt == 2:
            self.concat_fc = nn.Sequential(nn.LayerNorm(3 * self.hsz), nn.Dropout(p), nn.Linear(3 * self.hsz, self.hsz),
                                           nn.ReLU(True), nn.LayerNorm(self.hsz))
        self.input_embedding = nn.Sequential(nn.Dropout(p), nn.Linear(self.bridge_hsz, self.hsz), nn.ReLU(True),
                                             nn.LayerNorm(self.hsz))
        self.input_encoder = StackedEncoder(n_blocks=opt.input_encoder_n_blocks, n_conv=opt.input_encoder_n_conv,
                                            kernel_size=opt.input_encoder_kernel_size,
                                            num_heads=opt.input_encoder_n_heads, hidden_size=self.hsz, dropout=p)
        self.str_attn = StructuredAttention(dropout=p, scale=opt.scale, add_void=opt.add_non_visual)
        self.c2q_down_projection = nn.Sequential(nn.LayerNorm(3 * self.hsz), nn.Dropout(p),
                                                 nn.Linear(3 * self.hsz, self.hsz), nn.ReLU(True))
        self.cls_encoder = StackedEncoder(n_blocks=opt.cls_encoder_n_blocks, n_conv=opt.cls_encoder_n_conv,
                                          kernel_size=opt.cls_encoder_kernel_size, num_heads=opt.cls_encoder_n_heads,
                                          hidden_size=self.hsz, dropout=p)
        self.cls_projection_layers = nn.ModuleList(
            [LinearWrapper(self.hsz, self.hsz, layer_norm=True, dropout=p, relu=True)] +
            [ConvLinear(self.hsz, self.hsz, kernel_size=3, layer_norm=True, dropout=p, relu=True)
             for _ in range(self.t_iter)])
        self.temporal_scoring_st_layers = nn.ModuleList(
            [LinearWrapper(self.hsz, 1, layer_norm=True, dropout=p, relu=False) for _ in range(self.t_iter + 1)])
        self.temporal_scoring_ed_layers = nn.ModuleList(
            [LinearWrapper(self.hsz, 1, layer_norm=True, dropout=p, relu=False) for _ in range(self.t_iter + 1)])
        self.temporal_criterion = nn.CrossEntropyLoss(reduction="sum")
        self.classifier = LinearWrapper(self.hsz * 2 if self.add_local else self.hsz, 1, layer_norm=True, dropout=p,
                                        relu=False)

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, batch):
        if self.inference_mode:
            return self.forward_main(batch)
        out, att_loss, att_predictions, temporal_loss, temporal_predictions, _ = self.forward_main(batch)
        return out, att_loss, att_predictions, temporal_loss, temporal_predictions

    def forward_main(self, batch):
        """batch: the edict ``prepare_inputs`` builds (tvqa_dataset.py:631-688); see the reference docstring
        (model/stage.py:200-219).  Eval returns ``(out, att_loss, att_predictions, temporal_loss, softmax(t_scores),
        other_outputs)``; ``inference_mode`` returns the prediction dict (model/stage.py:297-312)."""
        if self.training and self.dropout > 0 and not torch.is_grad_enabled():
            raise NotImplementedError(
                "tvqaplus_b200: dropout > 0 in train mode runs through the training engine (tvqaplus_b200/train.py), which "
                "needs gradients enabled; under torch.no_grad() call model.eval() or set opt.dropout = 0")
        if self.flag_cnt == 0:
            raise NotImplementedError  # model/stage.py:291-292
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(batch)  # outputs carry an autograd graph whose backward is our kernels (train.py)
        with torch.no_grad():
            return self._forward_eval(batch)

    # The LayerNorm(300) that closes bert_word_encoding_fc / vid_fc is not applied by the first kernel: it is folded into
    # input_embedding's Linear (LN(h) W^T + b = rstd * ((h - c) (W*gamma)^T - (mean - c) * cs) + cb, the row statistics
    # are gathered by the second kernel while it streams h), which halves the epilogue of the 768->300 GEMM.
    def _text_fc(self, x):
        fc, emb = self.bert_word_encoding_fc, self.input_embedding
        img, cs, cb = packed_ln(fc[2], fc[0], 32)
        h = ops.linear_fwd(x, img, None, self.bridge_hsz, ln_in=(cs, cb), relu=True)
        img2, cs2, cb2 = packed_ln(emb[1], fc[4], 64)
        return ops.linear_fwd(h, img2, None, self.hsz, ln_in=(cs2, cb2), relu=True, ln_out=_ln_pair(emb[3]))

    def _vid_fc(self, x):
        fc, emb = self.vid_fc, self.input_embedding
        img, cs, cb = packed_ln(fc[2], fc[0], 32)
        h = ops.linear_fwd(x, img, None, self.bridge_hsz, ln_in=(cs, cb), relu=True,
                           l2norm_in=True)  # F.normalize of model/stage.py:256 fused in
        img2, cs2, cb2 = packed_ln(emb[1], fc[4], 64)
        return ops.linear_fwd(h, img2, None, self.hsz, ln_in=(cs2, cb2), relu=True, ln_out=_ln_pair(emb[3]))

    @classmethod
    def base_encoder(cls, data, data_mask, init_encoder, downsize_encoder, input_encoder):
        """Raw features -> (N, L, hsz) embeddings; ``init_encoder`` / ``downsize_encoder`` are the bound fused
        projections (``_text_fc`` / ``_vid_fc`` already contain ``input_embedding``)."""
        h = init_encoder(data)
        if downsize_encoder is not None:
            h = downsize_encoder(h)
        return input_encoder(h, data_mask)

    def _cqa_weights(self):
        c2q, cat = self.c2q_down_projection, self.concat_fc
        params = (c2q[0].weight, c2q[0].bias, c2q[2].weight, c2q[2].bias, cat[0].weight, cat[0].bias, cat[2].weight, cat[2].bias)
        tag = tuple((t.data_ptr(), t._version) for t in params) + (_packing.epoch(),)
        if getattr(self, "_cqa_tag", None) != tag:  # re-fold after an optimizer step / load_state_dict / .to()
            self._cqa_blob, self._cqa_tag = ops.cqa_prepare_weights(*params), tag
        return dict(wblob=self._cqa_blob, cat_out_ln_g=cat[4].weight, cat_out_ln_b=cat[4].bias)

    def _head_weights(self):
        w = {}
        for key, mod in (("proj", self.cls_projection_layers[0]), ("st", self.temporal_scoring_st_layers[0]),
                         ("ed", self.temporal_scoring_ed_layers[0]), ("cls", self.classifier)):
            w[key + "_ln_g"], w[key + "_ln_b"] = mod.conv[0].weight, mod.conv[0].bias
            w[key + "_w"], w[key + "_b"] = mod.conv[2].weight, mod.conv[2].bias
        return w

    @staticmethod
    def invalidate_weight_cache():
        """Re-pack every weight image at the next forward.  Needed only after in-place edits through ``.data`` (they do not
        move the parameter's version counter, which is what the caches are keyed on)."""
        _packing.invalidate_weight_cache()

    def _monitor(self, device):
        mon = getattr(self, "_err_monitor", None)
        if mon is None or mon.device != device:
            from .._lib import DeviceErrorMonitor
            mon = self._err_monitor = DeviceErrorMonitor(device)
        return mon

    def check_device_errors(self):
        """Wait for the last forward and raise if one of its kernels reported a timed-out barrier wait."""
        mon = getattr(self, "_err_monitor", None)
        if mon is not None:
            mon.poll(wait=True)

    def _forward_eval(self, batch):
        # Device error flag (bounded mbarrier waits): polled without synchronising — a time-out in forward k raises at the
        # latest at the start of forward k+1 (or right here when this call reads results back to the host anyway).
        mon = self._monitor(batch.qas_bert.device) if self.error_checks else None
        if mon is not None:
            mon.poll()
        out = self._forward_eval_impl(batch)
        if mon is not None:
            mon.post()
            if self.inference_mode:
                mon.poll(wait=True)  # get_att_prediction has synchronised already
        return out

    def _forward_eval_impl(self, batch):
        self.bsz = bsz = len(batch.qid)
        num_a, hsz = self.num_a, self.hsz
        qas_mask = batch.qas_mask.float()
        a_embed = self.base_encoder(batch.qas_bert.view(bsz * num_a, -1, self.wd_size), qas_mask.view(bsz * num_a, -1),
                                    self._text_fc, None, self.input_encoder)
        Lq = a_embed.shape[1]
        a_embed = a_embed.view(bsz, num_a, Lq, hsz)

        if self.flag_cnt == 1:
            return self._forward_single_stream(batch, a_embed, qas_mask)
        T, Ls = batch.sub_bert.shape[1:3]
        sub_mask = batch.sub_mask.float()
        # the sub-title encoder writes its output straight into the operand images of the fused CQA kernel's workspace
        # (no fp32 round trip, no conversion pass) whenever the sequence length allows it
        cqa_ws = ops.CqaWorkspace(bsz, T, Lq, a_embed.device)
        if self.input_encoder.can_write_cqa_workspace(Ls):
            self.input_encoder(self._text_fc(batch.sub_bert.view(bsz * T, Ls, -1)), sub_mask.view(bsz * T, Ls), cqa_workspace=cqa_ws)
            sub_embed = None
        else:
            sub_embed = self.base_encoder(batch.sub_bert.view(bsz * T, Ls, -1), sub_mask.view(bsz * T, Ls),
                                          self._text_fc, None, self.input_encoder).view(bsz, T, Ls, hsz)
        R = batch.vid.shape[2]
        vid_mask = batch.vid_mask.float()
        vid_embed = self.base_encoder(batch.vid.view(bsz * T, R, -1), vid_mask.view(bsz * T, R),
                                      self._vid_fc, None, self.input_encoder).view(bsz, T, R, hsz)

        need_raw = self.inference_mode or (self.training and self.use_sup_att)  # vid_raw_s feeds get_att_prediction / get_att_loss
        blocked = ops.blocked_fits(Lq)  # the statement goes to the cls encoder in the blocked tile layout (never seen outside)
        cq = ops.cqa_fused_fwd(a_embed, qas_mask.view(bsz, num_a, Lq), sub_embed, sub_mask, vid_embed, vid_mask,
                               self._cqa_weights(), self.scale, want=("vid_raw_s",) if need_raw else (), blocked=blocked,
                               workspace=cqa_ws)
        other_outputs = {"vid_raw_s": cq.get("vid_raw_s")}
        out, target, t_scores, t_probs = self.classfier_head_multi_proposal(
            cq["statement"], cq["statement_mask"], batch.target, batch.ts_label, batch.ts_label_mask,
            extra_span_length=self.extra_span_length, _blocked=blocked)
        other_outputs["temporal_scores"] = t_scores
        return self._finish(batch, out, target, t_scores, t_probs, other_outputs)

    def _finish(self, batch, out, target, t_scores, t_probs, other_outputs):
        """Return conventions of forward_main (model/stage.py:297-348)."""
        if self.training:  # ([out (N_new, 5), target (N_new,)], att_loss, att_predictions, temporal_loss, raw t_scores, ...)
            att_loss, att_predictions = 0, None
            if self.use_sup_att and self.vfeat_flag:  # model/stage.py:317-341
                cur_loss, att_predictions = self.get_att_loss(
                    other_outputs["vid_raw_s"], batch.att_labels, batch.target, batch.qas, qids=batch.qid, q_lens=batch.q_l,
                    vid_names=batch.vid_name, img_indices=batch.image_indices, boxes=batch.boxes, start_indices=batch.anno_st_idx,
                    num_negatives=self.num_negatives, use_hard_negatives=batch.use_hard_negatives, drop_topk=self.drop_topk)
                att_loss = att_loss + cur_loss
            return [out, target], att_loss, att_predictions, self._last_ts_loss, t_scores, other_outputs
        if self.inference_mode:
            return {"answer": out, "t_scores": t_probs,
                    "att_predictions": self.get_att_prediction(
                        scores=other_outputs["vid_raw_s"], object_vocab=batch.eval_object_word_ids, words=batch.qas,
                        vid_names=batch.vid_name, qids=batch.qid, img_indices=batch.image_indices, boxes=batch.boxes,
                        start_indices=batch.anno_st_idx) if self.vfeat_flag else None}
        temporal_loss = self._last_ts_loss if self._last_ts_loss is not None else self.get_ts_loss(t_scores, batch.ts_label, batch.target)
        return out, 0, None, temporal_loss, t_probs, other_outputs

    def _forward_train(self, batch):
        """Train-mode forward with gradients (main.py:53-61): ``StageTrainFunction`` runs the training engine's kernels and its
        backward hands every parameter its gradient, so ``loss.backward()`` / ``optimizer.step()`` of the reference's loop
        work unchanged.  Return convention of model/stage.py:345-348."""
        from ..train import StageTrainFunction, TrainEngine
        eng = getattr(self, "_train_engine", None)
        if eng is None:
            eng = self._train_engine = TrainEngine(self)
        self.bsz = len(batch.qid)
        offsets = None
        if self.use_sup_att and self.vfeat_flag and getattr(batch, "att_labels", None) is not None:
            pos, neg = self._att_pairs(batch.att_labels, batch.target, batch.anno_st_idx, self.num_negatives,
                                       batch.use_hard_negatives, self.drop_topk)
            N, Lq = batch.qas_bert.shape[0], batch.qas_bert.shape[2]
            offsets = eng.att_offsets_from_pairs(pos, neg, N, batch.vid.shape[1], Lq, batch.vid.shape[2])
        params = [p for _, p in self.named_parameters()]
        out, att_loss, temporal_loss, t_scores, target = StageTrainFunction.apply(eng, batch, offsets, *params)
        if offsets is None:
            att_loss = 0
        return [out, target], att_loss, None, temporal_loss, t_scores, {"temporal_scores": t_scores}

    def _forward_single_stream(self, batch, a_embed, qas_mask):
        """sub-only / vfeat-only models (model/stage.py:283-290): the statement is that stream's mixed tensor, its mask the
        stream's own.  Not a hot path (``run_main.sh`` always uses both streams): it runs through the module-level kernels
        (attention-only CQA launch + the 384 -> 128 projection), which materialise the (N,5,T,Lq,384) concatenation."""
        bsz, num_a, hsz = self.bsz, self.num_a, self.hsz
        Lq = a_embed.shape[2]
        if self.sub_flag:
            raw, mask, fc = batch.sub_bert, batch.sub_mask.float(), self._text_fc
        else:
            raw, mask, fc = batch.vid, batch.vid_mask.float(), self._vid_fc
        T, Lc = raw.shape[1:3]
        ctx = self.base_encoder(raw.view(bsz * T, Lc, -1), mask.view(bsz * T, Lc), fc, None, self.input_encoder).view(bsz, 1, T, Lc, hsz)
        statement, statement_mask, raw_s, _ = self.qa_ctx_attention(
            a_embed.view(bsz, num_a, 1, Lq, hsz), ctx, qas_mask.view(bsz, num_a, 1, Lq), mask.view(bsz, 1, T, Lc), None, None)
        other_outputs = {"vid_raw_s": raw_s if self.vfeat_flag else None}
        out, target, t_scores, t_probs = self.classfier_head_multi_proposal(
            statement, statement_mask.contiguous(), batch.target, batch.ts_label, batch.ts_label_mask,
            extra_span_length=self.extra_span_length)
        other_outputs["temporal_scores"] = t_scores
        return self._finish(batch, out, target, t_scores, t_probs, other_outputs)

    # ------------------------------------------------------------------------------------------------ heads
    def classfier_head_multi_proposal(self, statement, statement_mask, targets, ts_labels, ts_labels_mask,
                                      max_num_proposal=1, ce_prob_thd=0.01, iou_thd=0.5, extra_span_length=3, _blocked=False):
        """statement (N,5,Li,Lqa,D), statement_mask (N,5,Li,Lqa) -> answer scores (N,5), targets, masked span logits
        (N,5,Li,2) [+ their softmax over Li].  cls_encoder + masked max-pool is one kernel, everything after it another
        (eval branch of the reference, model/stage.py:484-537 with get_proposals :439-467)."""
        bsz, num_a, num_img, num_words = statement_mask.shape
        n_seq = bsz * num_a * num_img
        if _blocked:  # statement = flat blocked tile buffer of cqa_fused_fwd(blocked=True)
            max_statement = self.cls_encoder(statement, statement_mask.view(n_seq, num_words), pool=True,
                                             blocked_in=(n_seq, num_words))
        else:
            max_statement = self.cls_encoder(statement.view(n_seq, num_words, -1), statement_mask.view(n_seq, num_words),
                                             pool=True)
        # get_ts_loss (model/stage.py:539-555) rides in the same launch whenever the st/ed labels are there
        labels = None
        if isinstance(ts_labels, dict) and targets is not None:
            labels = (targets, ts_labels["st"], ts_labels["ed"])
        train = self.training and self.add_local
        if train and labels is None:
            raise RuntimeError("tvqaplus_b200: the train-mode proposal pass needs targets and ts_labels = {st, ed}")
        w = self._head_weights()
        head = ops.span_head_fwd(max_statement.view(bsz * num_a, num_img, -1), statement_mask, ts_labels_mask.float(), w,
                                 add_local=self.add_local, extra_span_length=extra_span_length, labels=labels, want_train=train)
        self._last_spans = head["spans"]
        self._last_ts_loss = head.get("ts_loss")
        if train:
            # train branch of get_proposals (model/stage.py:407-438): ground-truth span + predicted span with IoU >= iou_thd
            # (the kernel implements the reference's defaults max_num_proposal = 1, iou_thd = 0.5)
            if max_num_proposal != 1 or iou_thd != 0.5:
                raise NotImplementedError("tvqaplus_b200: train-mode proposals are built for max_num_proposal=1, iou_thd=0.5")
            prop = ops.train_proposals_fwd(head, labels, w, extra_span_length=extra_span_length, ce_prob_thd=ce_prob_thd)
            keep = prop["row_valid"].bool()  # N_new is data dependent: this is the one host sync of the train-mode forward
            return prop["answer_rows"][keep], prop["row_target"][keep], head["t_scores"], head["t_probs"]
        return head["answer"], targets, head["t_scores"], head["t_probs"]

    def residual_temporal_predictor(self, layer_idx, input_tensor):
        """(N, L, D) -> (x + proj(x), cat[st, ed] (N, L, 2)) — module-level form of model/stage.py:469-482."""
        x = input_tensor + self.cls_projection_layers[layer_idx](input_tensor)
        st = self.temporal_scoring_st_layers[layer_idx](x)
        ed = self.temporal_scoring_ed_layers[layer_idx](x)
        return x, torch.cat([st, ed], dim=2)

    def qa_ctx_attention(self, qa_embed, ctx_embed, qa_mask, ctx_mask, noun_mask, non_visual_vectors):
        """One stream of model/stage.py:365-387 on its own: (mixed, mixed_mask, raw_s, s_normalized)."""
        if noun_mask is not None or non_visual_vectors is not None:
            raise NotImplementedError("tvqaplus_b200: noun_mask / non_visual_vectors are not built")
        u_a, raw_s, s_mask, s_norm = self.str_attn(qa_embed, ctx_embed, qa_mask, ctx_mask)
        num_img = ctx_mask.shape[2]
        qa_rep = qa_embed.expand(-1, -1, num_img, -1, -1)
        c2q = self.c2q_down_projection
        img, cs, cb = packed_ln(c2q[2], c2q[0], 64)
        mixed = ops.linear_fwd(torch.cat([qa_rep, u_a, qa_rep * u_a], dim=-1), img, None, self.hsz, ln_in=(cs, cb),
                               relu=True)
        return mixed, (s_mask.sum(-1) != 0).float(), raw_s, s_norm

    def get_ts_loss(self, temporal_scores, ts_labels, answer_indices):
        """Mean of the start / end cross-entropies (sum over examples) on the ground-truth answer's row
        (model/stage.py:539-555)."""
        rows = torch.arange(len(answer_indices), dtype=torch.long, device=temporal_scores.device)
        ca = temporal_scores[rows, answer_indices.to(temporal_scores.device)]
        st = F.cross_entropy(ca[:, :, 0], ts_labels["st"].to(ca.device), reduction="sum")
        ed = F.cross_entropy(ca[:, :, 1], ts_labels["ed"].to(ca.device), reduction="sum")
        return (st + ed) / 2.0

    # ------------------------------------------------------------------------------------------------ attention ranking loss
    @classmethod
    def sample_negatives(cls, pred_score, pos_indices, neg_indices, num_negatives=2, use_hard_negatives=False,
                         negative_pool_size=0, num_hard=2, drop_topk=0):
        """Random negatives for the positives of one (image, word) pair (model/stage.py:556-607, the branch the reference's
        defaults always take: ``hard_negative_start`` equals ``n_epoch``, config.py:25,77).  Index tensors live on the host;
        the draw is ``torch.randint`` on the global CPU generator, like the reference's."""
        if use_hard_negatives:
            raise NotImplementedError("tvqaplus_b200: hard-negative sampling is not built (never active with the reference's "
                                      "defaults, and its top-k branch crashes in the reference when a word has few negatives)")
        sampled_pos = torch.cat([pos_indices] * num_negatives, dim=0)
        sampled_neg = neg_indices[torch.randint(low=0, high=len(neg_indices), size=(len(sampled_pos),), dtype=torch.long)]
        return sampled_pos, sampled_neg

    def _att_pairs(self, att_labels, target, start_indices, num_negatives, use_hard_negatives, drop_topk):
        """(batch, answer, image, word, region) rows of every sampled positive / negative - the index bookkeeping of
        model/stage.py:613-694 and sample_negatives' random branch (:556-607), on the host (it depends on the labels and the
        host RNG only).  The reference makes one ``torch.randint(high=#negatives of the word)`` call per (example, image, word)
        in loop order; exactly those calls are made here, in that order, on the same global CPU generator (so the sampled
        pairs are the reference's for a given seed) - everything around them (nonzero / unique / per-word slicing /
        concatenation, ~10 tiny tensor ops per word in the reference) is a handful of numpy operations per labelled image:
        tens of milliseconds per batch instead of hundreds."""
        if use_hard_negatives:
            raise NotImplementedError("tvqaplus_b200: hard-negative sampling is not built (never active with the reference's "
                                      "defaults, and its top-k branch crashes in the reference when a word has few negatives)")
        tgt = target.tolist() if torch.is_tensor(target) else list(target)
        cols, draws = [], []  # per labelled image: (b, ca, img, word[], pos_region[], zero_sorted, word-of-draw[]) and its draws
        for b in range(len(tgt)):
            ca, st = int(tgt[b]), start_indices[b]
            for i, lab in enumerate(att_labels[b]):
                lab = np.asarray(lab.cpu().numpy() if torch.is_tensor(lab) else lab)
                nz = lab != 0
                w_idx, r_idx = np.nonzero(nz)          # row-major: by word, then region (torch.nonzero order)
                if w_idx.size == 0:
                    continue
                words, first, cnt = np.unique(w_idx, return_index=True, return_counts=True)
                reps = cnt * num_negatives
                grp = np.repeat(np.arange(len(words)), reps)                       # group (= word) of every draw
                within = np.arange(reps.sum()) - np.repeat(np.cumsum(reps) - reps, reps)
                pos_sel = first[grp] + within % cnt[grp]                           # cat([positives of the word] * num_negatives)
                n_zero = (~nz).sum(1)
                if np.any(n_zero[words] == 0):
                    raise RuntimeError("get_att_loss: a labelled word has no negative region (the reference's randint fails too)")
                zero_sorted = np.argsort(nz, axis=1, kind="stable")                # the zero regions of every word, ascending, first
                cols.append((b, ca, st + i, w_idx[pos_sel], r_idx[pos_sel], zero_sorted, words[grp]))
                draws.append(torch.cat([torch.randint(low=0, high=int(h), size=(int(k),), dtype=torch.long)
                                        for h, k in zip(n_zero[words].tolist(), reps.tolist())]).numpy())
        if not cols:  # the reference's torch.cat of an empty list raises here as well
            raise RuntimeError("get_att_loss: no positive attention label in the batch")
        pos_rows, neg_rows = [], []
        for (b, ca, img, w, r, zero_sorted, wd), draw in zip(cols, draws):
            head = np.broadcast_to(np.array([b, ca, img], dtype=np.int64), (draw.size, 3))
            pos_rows.append(np.concatenate([head, w[:, None], r[:, None]], 1))
            neg_rows.append(np.concatenate([head, wd[:, None], zero_sorted[wd, draw][:, None]], 1))
        return torch.from_numpy(np.concatenate(pos_rows, 0).astype(np.int64)), torch.from_numpy(np.concatenate(neg_rows, 0).astype(np.int64))

    def get_att_loss(self, scores, att_labels, target, words, vid_names, qids, q_lens, img_indices, boxes, start_indices,
                     num_negatives=2, use_hard_negatives=False, drop_topk=0):
        """Ranking loss between labelled and sampled unlabelled regions of the ground-truth answer's words
        (model/stage.py:609-746).  ``scores`` = vid_raw_s (N, 5, Li, Lqa, Lr) on the device.  The pair lists are built on the
        host, the gather + loss + sum is one kernel pair (stage_att_loss_fwd): no per-element ``.item()`` round trips.
        Returns ``(att_loss, att_predictions)``; ``att_predictions`` (eval mode only) as the reference builds it."""
        pos, neg = self._att_pairs(att_labels, target, start_indices, num_negatives, use_hard_negatives, drop_topk)
        strides = torch.tensor(scores.stride(), dtype=torch.long)
        if not scores.is_contiguous():
            raise RuntimeError("tvqaplus_b200: scores must be contiguous")
        att_loss = ops.att_loss_fwd(scores, (pos * strides).sum(1).to(scores.device), (neg * strides).sum(1).to(scores.device),
                                    loss_type=self.att_loss_type, alpha=self.alpha, margin=self.margin)
        att_predictions = None
        if not self.training and self.vfeat_flag:
            att_predictions = dict(det_q=[], det_ca=[])
            uniq = np.unique(pos.numpy(), axis=0)
            rows = scores[uniq[:, 0], uniq[:, 1], uniq[:, 2], uniq[:, 3]].cpu()  # ONE gather + device->host copy
            for row, sc in zip(uniq, rows):
                b, ca, img, wd, _ = [int(v) for v in row]
                st = start_indices[b]
                n_reg = att_labels[b][img - st].shape[1]
                if boxes is not None and n_reg != len(boxes[b][img - st]):
                    raise AssertionError("scores / boxes length mismatch for %s" % (row,))
                det = {"pred": sc[:n_reg].clone(), "word": words[b, ca, wd], "qid": qids[b], "vid_name": vid_names[b],
                       "img_idx": img_indices[b][img], "boxes": boxes[b][img - st] if boxes is not None else None}
                (att_predictions["det_q"] if wd < q_lens[b] else att_predictions["det_ca"]).append(det)
        return att_loss, att_predictions

    # ------------------------------------------------------------------------------------------------ inference post-processing
    def get_att_prediction(self, scores, object_vocab, words, vid_names, qids, img_indices, boxes, start_indices,
                           score_thd=0.2):
        """Per (question, answer, annotated frame, object word): regions whose cosine score >= ``score_thd``, ascending
        (model/stage.py:748-806).  ONE device->host copy of ``scores`` instead of one per word."""
        if not self.vfeat_flag or scores is None or boxes is None:
            return None
        s = scores.detach().cpu().numpy()
        w = words.cpu().numpy()
        vocab = set(int(v) for v in object_vocab)
        preds = []
        for b in range(len(s)):
            st0 = start_indices[b]
            per_q = {}
            for a in range(self.num_a):
                cur = []
                obj_pos = [i for i, wid in enumerate(w[b, a].tolist()) if wid in vocab]
                for local_i, img_boxes in enumerate(boxes[b]):
                    gi = local_i + st0
                    for wi in obj_pos:
                        row = s[b, a, gi, wi]
                        keep = np.nonzero(row >= score_thd)[0]
                        order = keep[np.argsort(row[keep])]
                        cur.append({"pred": [float(row[i]) for i in order], "bbox": [img_boxes[i] for i in order],
                                    "word": int(w[b, a, wi]), "qid": int(qids[b]), "vid_name": vid_names[b],
                                    "img_idx": img_indices[b][gi]})
                per_q[a] = cur
            preds.append(per_q)
        return preds

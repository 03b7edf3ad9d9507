"""AdjustLabelSmoothedCrossEntropyCriterion on the fused loss kernel.

Mirrors criterions/label_smoothed_cross_entropy.py:129-275 (same constructor arguments, `forward(model, sample,
update_num, reduce)` -> `(loss, sample_size, logging_output)` with the same logging keys, multi-task list recursion
of :175-202, R-Drop sample duplication of :56-71 incl. the doubled `sample_patch_num`)."""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import ops
from ._fairseq_compat import FairseqCriterion, HAVE_FAIRSEQ, metrics

try:        # the criterion's flags are a fairseq dataclass in the reference (:14-53); fairseq derives --label-smoothing etc. from it
    from fairseq.dataclass import FairseqDataclass
    from omegaconf import II
    _SENTENCE_AVG = II("optimization.sentence_avg")
except ImportError:
    FairseqDataclass, _SENTENCE_AVG = object, False


@dataclass
class AdjustLabelSmoothedCrossEntropyCriterionConfig(FairseqDataclass):
    label_smoothing: float = field(default=0.0, metadata={"help": "epsilon for label smoothing, 0 means no label smoothing"})
    report_accuracy: bool = field(default=False, metadata={"help": "report accuracy metric"})
    ignore_prefix_size: int = field(default=0, metadata={"help": "Ignore first N tokens"})
    ignore_eos: bool = field(default=False, metadata={"help": "Ignore eos token"})
    sentence_avg: bool = _SENTENCE_AVG
    drop_worst_ratio: float = field(default=0.0, metadata={"help": "ratio for discarding bad samples"})
    drop_worst_after: int = field(default=0, metadata={"help": "steps for discarding bad samples"})
    use_rdrop: bool = field(default=False, metadata={"help": "use R-Drop"})
    reg_alpha: float = field(default=1.0, metadata={"help": "weight for R-Drop"})
    sample_patch_num: int = field(default=196, metadata={"help": "sample patches for v1"})
    constraint_range: Optional[str] = field(default=None, metadata={"help": "constraint range"})


def construct_rdrop_sample(x):
    if isinstance(x, dict):
        return {k: construct_rdrop_sample(v) for k, v in x.items()}
    if isinstance(x, torch.Tensor):
        if x.dim() == 0:            # a collater count living on the device (CUDA-graph path): doubled like the ints below
            return x * 2
        return x.repeat(2, *([1] * (x.dim() - 1)))
    if isinstance(x, bool) or x is None:
        return x
    if isinstance(x, int):
        return x * 2
    if isinstance(x, np.ndarray):
        return x.repeat(2)
    raise NotImplementedError(type(x))


class AdjustLabelSmoothedCrossEntropyCriterion(FairseqCriterion):
    """A FairseqCriterion (fairseq's `register_criterion` accepts nothing else); `batch_task_*` are this package's own
    switches for the multi-task batching and default to on."""

    def __init__(self, task, sentence_avg=False, label_smoothing=0.0, ignore_prefix_size=0, ignore_eos=False,
                 report_accuracy=False, drop_worst_ratio=0, drop_worst_after=0, use_rdrop=False, reg_alpha=1.0,
                 sample_patch_num=196, constraint_range=None, batch_task_stems=True, batch_task_encoders=True,
                 batch_task_decoders=True):
        super().__init__(task)
        if report_accuracy:
            raise NotImplementedError("--report-accuracy crashes in the reference (compute_accuracy unpacks 2 of 3 values, "
                                      "SURVEY.md appendix C) and no script sets it")
        self.batch_task_stems = batch_task_stems
        self.batch_task_encoders = batch_task_encoders and batch_task_stems
        self.batch_task_decoders = batch_task_decoders and self.batch_task_encoders
        self.task = task
        self.padding_idx = task.target_dictionary.pad()
        self.eos_idx = task.target_dictionary.eos()
        self.sentence_avg = sentence_avg
        self.eps = label_smoothing
        self.ignore_prefix_size, self.ignore_eos = ignore_prefix_size, ignore_eos
        self.report_accuracy = report_accuracy
        self.drop_worst_ratio, self.drop_worst_after = drop_worst_ratio, drop_worst_after
        self.use_rdrop, self.reg_alpha = use_rdrop, reg_alpha
        self.sample_patch_num = sample_patch_num
        self.constraint_range = None
        if constraint_range is not None:
            cs, ce = constraint_range.split(",")
            self.constraint_range = (int(cs), int(ce))

    def _batch_stems(self, model, sample):
        """Multi-task micro-step: push the images of all image tasks through the ResNet stem in ONE grouped pass (BatchNorm
        statistics stay per task, running statistics are updated in task order: musketeer_b200/resnet.py), and hand every
        task its slice as `patch_features`.  Arithmetic per task is unchanged; every stem kernel runs once instead of
        once per task.  Skipped when the tasks' image batches differ in shape or with ResNet drop-path.

        The flags of run_scripts/musketeer/train_musketeer.sh keep the merged passes: with R-Drop every image task's sample is
        duplicated HERE (construct_rdrop_sample, label_smoothed_cross_entropy.py:56-71,204-207) before the grouped stem, so a
        task's group is what its own forward would have pushed through the stem; dropout / drop-path draw independent
        per-element / per-row masks, which is what the per-task passes do too (other random numbers, the same distribution);
        patch sampling (`sample_patch_num`, :175-178: every task but the last of the list) is applied inside the merged pass,
        the per-row patch subsets drawn in task order like the per-task passes draw them."""
        enc = getattr(model, "encoder", None)
        stem = getattr(enc, "embed_images", None)
        if stem is None or not self.batch_task_stems:
            return sample
        idx = [i for i, s in enumerate(sample) if s["net_input"].get("patch_images") is not None
               and s["net_input"].get("patch_features") is None]
        if len(idx) < 2 or len(idx) > 8:
            return sample
        imgs = [sample[i]["net_input"]["patch_images"] for i in idx]
        if any(im.shape != imgs[0].shape or im.dtype != imgs[0].dtype for im in imgs):
            return sample
        if getattr(stem, "drop_path_rate", 0.0) > 0.0 and stem.training:
            return sample
        if self.use_rdrop:
            sample = list(sample)
            for i in idx:
                sample[i] = dict(construct_rdrop_sample(sample[i]), _rdrop_done=True)
            imgs = [sample[i]["net_input"]["patch_images"] for i in idx]
        b = imgs[0].shape[0]
        feats_all = stem(torch.cat(imgs, 0), groups=len(idx))
        hw = stem.last_hw
        out = list(sample)
        nis = [sample[i]["net_input"] for i in idx]
        sampled = self.sample_patch_num > 0
        merge = (self.batch_task_encoders and
                 all(ni.get("sample_patch_num") is None and ni.get("patch_images_2") is None for ni in nis) and
                 # the recursion below sets sample_patch_num on every task but the LAST of the list (:175-178): the merged pass
                 # samples patches for all of its tasks or for none
                 (not sampled or all(i < len(sample) - 1 for i in idx)))
        if merge:
            # ONE encoder pass for all image tasks: source tokens right-padded to the longest prompt (padded keys are masked
            # and padded rows zeroed, so every real position computes what its own task's pass computes), then each task
            # decodes against its slice of the encoder output.  Linear / LayerNorm / FFN kernels see 4x the rows.
            pad = enc.padding_idx
            S = max(ni["src_tokens"].shape[1] for ni in nis)
            src = torch.cat([torch.nn.functional.pad(ni["src_tokens"], (0, S - ni["src_tokens"].shape[1]), value=pad)
                             for ni in nis], 0)
            masks = torch.cat([ni["patch_masks"] for ni in nis], 0)
            eo = enc(src, src_lengths=None, patch_images=imgs[0], patch_masks=masks, patch_features=(feats_all, hw),
                     # (construct_rdrop_sample doubles every int of the sample, sample_patch_num included -- the per-task
                     # path and the reference, label_smoothed_cross_entropy.py:61-62, keep 2 x sample_patch_num patches
                     # under R-Drop; so does the merged pass)
                     sample_patch_num=self.sample_patch_num * (2 if self.use_rdrop else 1) if sampled else None)
            xs = eo["encoder_out"][0].transpose(0, 1).split(b, 0)              # [b, N, d] per task; split: one cat backward
            pms = eo["encoder_padding_mask"][0].split(b, 0)
            pos = eo["position_embeddings"][0].split(b, 0)
            for k, i in enumerate(idx):
                s = dict(sample[i])
                s["net_input"] = dict(s["net_input"])
                s["net_input"]["encoder_out"] = {"encoder_out": [xs[k].transpose(0, 1)], "encoder_padding_mask": [pms[k]],
                                                 "position_embeddings": [pos[k]], "encoder_embedding": [],
                                                 "encoder_states": [], "src_tokens": [], "src_lengths": []}
                out[i] = s
            if self.batch_task_decoders and not self.use_rdrop:      # (the R-Drop KL term pairs the two halves of ONE task's logits)
                xs, pms, pos, idx2 = list(xs), list(pms), list(pos), list(idx)
                # text-only tasks with the same batch size join the decoder groups: their own encoder pass, right-padded
                # (masked) to the merged pass's source length
                N = xs[0].shape[1]
                for i, smp in enumerate(sample):
                    ni = smp["net_input"]
                    if (i in idx or ni.get("patch_images") is not None or ni.get("encoder_out") is not None
                            or ni["src_tokens"].shape[0] != b or ni["src_tokens"].shape[1] > N):
                        continue
                    e1 = enc(ni["src_tokens"], src_lengths=ni.get("src_lengths"))
                    x1, pm1, ps1 = e1["encoder_out"][0].transpose(0, 1), e1["encoder_padding_mask"][0], e1["position_embeddings"][0]
                    padn = N - x1.shape[1]
                    xs.append(torch.nn.functional.pad(x1, (0, 0, 0, padn)))
                    pms.append(torch.nn.functional.pad(pm1, (0, padn), value=True))
                    pos.append(torch.nn.functional.pad(ps1, (0, 0, 0, padn)))
                    s1 = dict(smp)
                    s1["net_input"] = dict(ni)
                    s1["net_input"]["encoder_out"] = {"encoder_out": [xs[-1].transpose(0, 1)], "encoder_padding_mask": [pms[-1]],
                                                      "position_embeddings": [pos[-1]], "encoder_embedding": [],
                                                      "encoder_states": [], "src_tokens": [], "src_lengths": []}
                    out[i] = s1
                    idx2.append(i)
                self._batch_decoders(model, out, idx2, xs, pms, pos, b)
            return out
        feats = feats_all.split(b, 0)                                           # split: one cat in the backward
        for k, i in enumerate(idx):
            s = dict(sample[i])
            s["net_input"] = dict(s["net_input"])
            s["net_input"]["patch_features"] = (feats[k], hw)
            out[i] = s
        return out

    def _batch_decoders(self, model, out, idx, xs, pms, pos, b):
        """Tasks of the merged encoder pass whose targets have similar lengths share ONE decoder pass + ONE loss launch
        (targets right-padded to the longest: padded positions are masked keys and pad targets).  The per-row losses come
        back as a vector, so every task still gets the sum over its own rows and its own sample size.  Tasks with
        constraint masks / confidences keep their own pass."""
        pad = self.padding_idx
        cand = [(k, i) for k, i in enumerate(idx) if out[i].get("constraint_masks") is None and out[i].get("conf") is None]
        cand.sort(key=lambda ki: out[ki[1]]["net_input"]["prev_output_tokens"].shape[1])
        groups, cur = [], []
        for k, i in cand:
            T = out[i]["net_input"]["prev_output_tokens"].shape[1]
            if cur:
                T0 = out[cur[0][1]]["net_input"]["prev_output_tokens"].shape[1]      # shortest of the group
                if not (T <= 32 or T0 >= 0.8 * T):
                    groups.append(cur)
                    cur = []
            cur.append((k, i))
        if cur:
            groups.append(cur)
        for grp in groups:
            if len(grp) < 2:
                continue
            Tm = max(out[i]["net_input"]["prev_output_tokens"].shape[1] for _, i in grp)
            prevs, tgts = [], []
            for _, i in grp:
                prev, tgt = out[i]["net_input"]["prev_output_tokens"], out[i]["target"]
                if self.ignore_prefix_size > 0:
                    tgt = tgt.clone()
                    tgt[:, :self.ignore_prefix_size] = pad
                if self.ignore_eos:
                    tgt = tgt.masked_fill(tgt.eq(self.eos_idx), pad)
                prevs.append(torch.nn.functional.pad(prev, (0, Tm - prev.shape[1]), value=pad))
                tgts.append(torch.nn.functional.pad(tgt, (0, Tm - tgt.shape[1]), value=pad))
            eo = {"encoder_out": [torch.cat([xs[k] for k, _ in grp], 0).transpose(0, 1)],
                  "encoder_padding_mask": [torch.cat([pms[k] for k, _ in grp], 0)],
                  "position_embeddings": [torch.cat([pos[k] for k, _ in grp], 0)]}
            logits, _ = model.decoder(torch.cat(prevs, 0), encoder_out=eo, padded_logits=True)
            if hasattr(model, "dec_timer"):
                model.dec_timer[1] += 1
            n = b * Tm
            # every row's loss will be divided by its task's sample size (forward(): loss_v1 / ss1 + ...): promise that
            # upstream gradient to the loss kernel, so the logits-sized gradient is written once, already scaled
            expected = None
            if self.ignore_prefix_size == 0 and not self.ignore_eos and all("ntokens" in out[i] for _, i in grp):
                ones = torch.ones(n, dtype=torch.float32, device=logits.device)
                expected = torch.cat([ones / (out[i]["target"].size(0) if self.sentence_avg else out[i]["ntokens"])
                                      for _, i in grp])
            loss_rows, nll_rows = ops.ls_cross_entropy_rows(logits, torch.cat(tgts, 0), self.eps, pad,
                                                            crange=self.constraint_range, expected_grad=expected)
            for g, (_, i) in enumerate(grp):
                out[i]["_precomputed"] = (loss_rows[g * n:(g + 1) * n].sum(), nll_rows[g * n:(g + 1) * n].sum())

    def forward(self, model, sample, update_num=0, reduce=True, _top=True):
        if isinstance(sample, list) and len(sample) > 1 and _top:
            sample = self._batch_stems(model, sample)
        if isinstance(sample, list) and len(sample) > 1:                               # :175-202
            if self.sample_patch_num > 0:
                sample[0]["net_input"]["sample_patch_num"] = self.sample_patch_num
            loss_v1, ss1, log1 = self.forward(model, sample[0], update_num, reduce, _top=False)
            loss_v2, ss2, log2 = self.forward(model, sample[1:], update_num, reduce, _top=False)
            loss = loss_v1 / ss1 + loss_v2 / ss2
            logging_output = {
                "loss": loss.data, "loss_v1": loss_v1.data, "loss_v2": loss_v2.data,
                "nll_loss": log1["nll_loss"].data / ss1 + log2["nll_loss"].data / ss2,
                "ntokens": log1["ntokens"] + log2["ntokens"],
                "nsentences": log1["nsentences"] + log2["nsentences"],
                "sample_size": 1, "sample_size_v1": ss1, "sample_size_v2": ss2,
            }
            return loss, 1, logging_output
        sample = sample[0] if isinstance(sample, list) else sample
        if self.use_rdrop and not sample.get("_rdrop_done"):
            sample = construct_rdrop_sample(sample)
        drop = self.drop_worst_ratio if (self.drop_worst_ratio > 0 and update_num > self.drop_worst_after) else 0.0
        pre = sample.get("_precomputed") if drop == 0 else None      # this task's rows of a merged decoder pass
        if pre is None:
            logits, _ = model(**sample["net_input"], padded_logits=True)
        target = sample["target"]
        if self.ignore_prefix_size > 0:                                                 # :239-243
            target = target.clone()
            target[:, :self.ignore_prefix_size] = self.padding_idx
        if self.ignore_eos:                                                             # :244-250
            target = target.masked_fill(target.eq(self.eos_idx), self.padding_idx)
        if pre is not None:
            loss, nll_rows = pre
        else:
            loss, nll_rows = ops.ls_cross_entropy(
                logits, target, self.eps, self.padding_idx, cmask=sample.get("constraint_masks"), conf=sample.get("conf"),
                crange=self.constraint_range, rdrop=self.use_rdrop, reg_alpha=self.reg_alpha, drop_worst_ratio=drop)
        if drop > 0:          # ntokens = rows kept (label_smoothed_cross_entropy.py:113)
            n = int(target.ne(self.padding_idx).sum())
            ntokens = 2 * int((n // 2) * (1 - drop)) if self.use_rdrop else int(n * (1 - drop))
        elif self.ignore_prefix_size > 0 or self.ignore_eos or "ntokens" not in sample:
            ntokens = int(target.ne(self.padding_idx).sum())   # host sync, like the reference's boolean indexing (:258-260)
        else:
            ntokens = sample["ntokens"]    # the collater's count of non-pad target tokens: same number, no device sync
        sample_size = sample["target"].size(0) if self.sentence_avg else ntokens
        logging_output = {"loss": loss.data, "nll_loss": nll_rows.sum().data, "ntokens": sample["ntokens"],
                          "nsentences": sample["nsentences"], "sample_size": sample_size}
        return loss, sample_size, logging_output

    @classmethod
    def reduce_metrics(cls, logging_outputs) -> None:
        """Aggregate the logging outputs of the data-parallel workers (label_smoothed_cross_entropy.py:286-345)."""
        if metrics is None:
            raise RuntimeError("reduce_metrics needs fairseq.metrics")
        tot = lambda key: sum(log.get(key, 0) for log in logging_outputs)
        sample_size, ntokens = tot("sample_size"), tot("ntokens")
        ss1, ss2 = max(tot("sample_size_v1"), 1), max(tot("sample_size_v2"), 1)
        metrics.log_scalar("loss", tot("loss") / sample_size, sample_size, round=3)
        metrics.log_scalar("loss_v1", tot("loss_v1") / ss1, ss1, round=3)
        metrics.log_scalar("loss_v2", tot("loss_v2") / ss2, ss2, round=3)
        metrics.log_scalar("nll_loss", tot("nll_loss") / sample_size, ntokens, round=3)
        try:
            from fairseq import utils as fs_utils
            metrics.log_derived("ppl", lambda meters: fs_utils.get_perplexity(meters["nll_loss"].avg))
        except (ImportError, AttributeError):
            pass
        metrics.log_scalar("ntokens", ntokens, 1, round=3)
        metrics.log_scalar("nsentences", tot("nsentences"), 1, round=3)
        metrics.log_scalar("sample_size", sample_size, 1, round=3)
        metrics.log_scalar("sample_size_v1", tot("sample_size_v1"), 1, round=3)
        metrics.log_scalar("sample_size_v2", tot("sample_size_v2"), 1, round=3)

    @staticmethod
    def logging_outputs_can_be_summed():
        return True

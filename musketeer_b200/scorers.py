"""All-candidate inference: every sentence of a batch is scored against every answer of a closed candidate set
and the best one is taken -- tasks/mm_tasks/vqa_gen.py:257-306 (`val_inference_type == "allcand"`), tasks/mm_tasks/snli_ve.py:
171-215 and utils/eval_utils.py:161-217 (eval_vqa_gen) / :249-312 (eval_snli_ve), which all share one body.

What the reference does per chunk of `valid_batch_size` answers: build prev / target / a DENSE bool constraint mask
[B*C, T, V] on the host, repeat_interleave the encoder output C times, run the decoder, masked_fill + full-vocabulary
log_softmax + gather.  Here, per chunk:
  * the encoder output is NOT replicated: the C candidate rows of a sentence are folded into the query axis of the
    cross-attention (ofa.TransformerDecoderLayer), and the per-layer cross K / V projections are computed once per batch and
    reused by every chunk (`_cross_kv_memo`);
  * the vocabulary projection runs only on the positions that count (answer tokens + eos, not the prompt or the padding);
  * the constraint is the CSR trie on the device: a position reads just the logits of its node's children
    (csrc/beam.cu trie_score_kernel), no mask tensor exists.
Scores equal the reference's `valid_result` [B, n_answers] (fp32 mode: 1e-4; tests/test_model_gpu.py)."""
from typing import List, Sequence

import torch

from . import ops
from .sequence_generator import flatten_trie


class AllCandidateScorer:
    """answers: list of 1-D int64 token tensors (no bos / eos), in `index2ans` order (vqa_gen.py:161-167);
    constraint_trie: the task's Trie holding [bos] + answer + [eos] for every answer (vqa_gen.py:158-167), or None for
    unconstrained scoring; valid_batch_size: answers per decoder pass (vqa_gen.py:181-183)."""

    def __init__(self, model, answers: Sequence[torch.Tensor], constraint_trie=None, valid_batch_size: int = 20,
                 bos: int = 0, pad: int = 1, eos: int = 2):
        self.model = model
        self.answers = [a.long().cpu() for a in answers]
        self.trie = constraint_trie
        self.valid_batch_size = int(valid_batch_size)
        self.bos, self.pad, self.eos = bos, pad, eos
        self._csr = None
        # trie node of every counted position of every answer: next layer after [bos] + answer[:i], i = 0 .. len(answer)
        self._nodes: List[List[int]] = []
        if constraint_trie is not None:
            self._csr_host, index = flatten_trie(constraint_trie, "cpu", return_index=True)
            kids = lambda n: n.child if hasattr(n, "child") else n
            for a in self.answers:
                cur, path = kids(constraint_trie.root).get(bos), []
                for i in range(len(a) + 1):
                    if cur is None:
                        raise ValueError("answer %s is not in the constraint trie" % a.tolist())
                    path.append(index[id(cur)])
                    if i < len(a):
                        cur = kids(cur).get(int(a[i]))
                self._nodes.append(path)
        else:
            self._nodes = [[-1] * (len(a) + 1) for a in self.answers]

    @classmethod
    def from_task(cls, task, model, valid_batch_size=None):
        """Built from the fields the reference tasks hold after build_model (vqa_gen.py:158-183): valid_answers_list (chunks),
        constraint_trie, src_dict."""
        answers = [a for chunk in task.valid_answers_list for a in chunk]
        vb = valid_batch_size or len(task.valid_answers_list[0])
        d = task.src_dict
        return cls(model, answers, task.constraint_trie, vb, bos=d.bos(), pad=d.pad(), eos=d.eos())

    @torch.no_grad()
    def score(self, sample):
        """-> valid_result [B, n_answers] fp32 (eval_utils.py:212)."""
        model = self.model
        ni = sample["net_input"]
        dev = ni["src_tokens"].device
        enc = model.encoder(ni["src_tokens"], src_lengths=ni.get("src_lengths"), patch_images=ni.get("patch_images"),
                            patch_masks=ni.get("patch_masks"))
        enc = dict(enc)
        enc["_cross_kv_memo"] = {}
        if self.trie is not None and (self._csr is None or self._csr[0].device != dev):
            self._csr = tuple(t.to(dev) for t in self._csr_host)
        prompts = [list(map(int, p)) for p in sample["decoder_prompts"]]
        B = len(prompts)
        out = []
        for c0 in range(0, len(self.answers), self.valid_batch_size):
            ans = self.answers[c0:c0 + self.valid_batch_size]
            C = len(ans)
            T = max(len(p) for p in prompts) + max(len(a) for a in ans)
            prev = torch.full((B * C, T), self.pad, dtype=torch.long)
            sel, node, tgt, seg = [], [], [], [0]
            for b, p in enumerate(prompts):
                for c, a in enumerate(ans):
                    r = b * C + c
                    prev[r, :len(p)] = torch.tensor(p, dtype=torch.long)
                    prev[r, len(p):len(p) + len(a)] = a
                    # counted positions: len(p) - 1 + i, i = 0 .. len(a): target answer[i] (eos last), node of [bos] + answer[:i]
                    for i in range(len(a) + 1):
                        sel.append(r * T + len(p) - 1 + i)
                        node.append(self._nodes[c0 + c][i])
                        tgt.append(int(a[i]) if i < len(a) else self.eos)
                    seg.append(len(sel))
            feats, _ = model.decoder(prev.to(dev), encoder_out=enc, features_only=True)
            rows = feats.reshape(B * C * T, -1).index_select(0, torch.tensor(sel, dtype=torch.long, device=dev))
            logits = model.decoder.output_layer(rows, padded=True)
            sc = ops.trie_score(logits, torch.tensor(seg, dtype=torch.int32, device=dev),
                                torch.tensor(node, dtype=torch.int32, device=dev), torch.tensor(tgt, dtype=torch.long, device=dev),
                                self._csr, self.pad)
            out.append(sc.view(B, C))
        return torch.cat(out, dim=-1)

    def predict(self, sample):
        """-> index of the best answer per sentence (eval_utils.py:213)."""
        return self.score(sample).argmax(1).tolist()

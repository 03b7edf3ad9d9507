"""Beam search over the incremental OFA decoder (models/sequence_generator.py:19-764 + fairseq BeamSearch.step as
spelled out in models/search.py:109-144).  Same constructor arguments, `generate(models, sample, **kw)` signature and
result layout (list over sentences of hypothesis dicts sorted by score: tokens / score / attention / alignment /
positional_scores).  Per step, on the kernels of libofa_b200.so: one token per beam through the decoder (paged self-attention
KV cache, cross-attention K / V projected once per sentence, csrc/decode.cu), then ONE fused launch pair for the tail --
temperature, constraint range / constraint trie, fp32 log-softmax, min-len / max-len / pad / unk masks, n-gram blocking,
previous beam scores and the top 2*beam selection (csrc/beam.cu).  The constraint trie (utils/trie.py; built by the tasks,
tasks/mm_tasks/vqa_gen.py:158-167) is flattened to CSR once and walked on the device; a step in which no hypothesis ends
does its beam bookkeeping in one launch (beam_advance_kernel), a step with finished hypotheses follows the reference's own
sequence of index operations.  Forced prefixes (`prefix_tokens`, the decoder prompt of the VQA beam-search evaluation:
tasks/mm_tasks/vqa_gen.py:311, utils/eval_utils.py:152) are applied inside the same fused tail.

Out of scope here (raises): lexical constraints, image-code / box generation, LM fusion, ensembles > 1."""
import math
from typing import Dict, List, Optional

import torch

from . import ops


def flatten_trie(trie, device, return_index=False):
    """utils/trie.py Trie -> CSR tensors (ptr [n+1], tok [e], child [e]); the root is node 0.  Children are stored in
    insertion order, like `list(node.child.keys())`.  Nodes are TreeNode objects (`.child` dict) or plain nested dicts.
    return_index=True also returns {id(node object): CSR node index}."""
    nodes, ptr, tok, child = [trie.root], [0], [], []
    i = 0
    while i < len(nodes):
        kids = nodes[i].child if hasattr(nodes[i], "child") else nodes[i]
        for t, c in kids.items():
            tok.append(int(t))
            child.append(len(nodes))
            nodes.append(c)
        ptr.append(len(tok))
        i += 1
    mk = lambda v: torch.tensor(v if v else [0], dtype=torch.int32, device=device)
    csr = (mk(ptr), mk(tok), mk(child))
    if return_index:
        return csr, {id(n): k for k, n in enumerate(nodes)}
    return csr


class SequenceGenerator(torch.nn.Module):
    def __init__(self, models, tgt_dict, beam_size=1, max_len_a=0, max_len_b=200, max_len=0, min_len=1,
                 normalize_scores=True, len_penalty=1.0, unk_penalty=0.0, temperature=1.0, match_source_len=False,
                 no_repeat_ngram_size=0, search_strategy=None, eos=None, symbols_to_strip_from_output=None,
                 lm_model=None, lm_weight=1.0, constraint_trie=None, constraint_range=None, gen_code=False,
                 gen_box=False, ignore_eos=False, zero_shot=False, cuda_graphs=True):
        super().__init__()
        models = list(models) if isinstance(models, (list, tuple)) else [models]
        if len(models) != 1 or lm_model is not None:
            raise NotImplementedError("ensembles / LM fusion are outside the hot-path scope")
        if gen_code or gen_box or search_strategy is not None or match_source_len:
            raise NotImplementedError("code/box generation and custom search strategies are outside the hot-path scope")
        if beam_size > 8:
            raise NotImplementedError("beam sizes above 8 (the beams of a sentence share one cross-attention cache row group)")
        self.model = models[0]
        self.tgt_dict = tgt_dict
        self.pad, self.unk, self.bos = tgt_dict.pad(), tgt_dict.unk(), tgt_dict.bos()
        self.eos = tgt_dict.eos() if eos is None else eos
        self.vocab_size = len(tgt_dict)
        self.beam_size = min(beam_size, self.vocab_size - 1)
        self.max_len_a, self.max_len_b, self.min_len = max_len_a, max_len_b, min_len
        self.max_len = max_len or self.model.max_decoder_positions()
        self.normalize_scores, self.len_penalty, self.unk_penalty = normalize_scores, len_penalty, unk_penalty
        self.temperature = temperature
        self.no_repeat_ngram_size = no_repeat_ngram_size
        self.ignore_eos = ignore_eos
        self.zero_shot = zero_shot            # constraints are applied AFTER the softmax (sequence_generator.py:878-889)
        self.constraint_trie = constraint_trie
        self._trie_csr = None
        self.constraint_start = self.constraint_end = None
        if constraint_range is not None:
            cs, ce = constraint_range.split(",")
            self.constraint_start, self.constraint_end = int(cs), int(ce)
        assert temperature > 0, "--temperature must be greater than 0"
        self.model.eval()
        # Decoder steps >= 1 are captured as CUDA graphs per (rows, source length, step) on the second call with a given shape
        # and replayed afterwards while no sentence has finished (the batch shrinks then and the search continues eagerly):
        # a step is ~300 small launches, so the beam search is host-bound without it.
        self.cuda_graphs = cuda_graphs
        self._static = {}

    @torch.no_grad()
    def forward(self, sample, prefix_tokens=None, bos_token=None):
        return self._generate([self.model], sample, prefix_tokens, bos_token=bos_token)

    @torch.no_grad()
    def generate(self, models, sample, **kwargs):
        return self._generate(models, sample, **kwargs)

    def _generate(self, models, sample, prefix_tokens=None, constraints=None, bos_token=None):
        if constraints is not None:
            raise NotImplementedError("lexical constraints are outside the hot-path scope")
        if prefix_tokens is not None and self.constraint_start is not None:
            raise NotImplementedError("prefix tokens together with a constraint range (no caller in the reference)")
        model = models[0] if isinstance(models, (list, tuple)) else models
        net_input = sample["net_input"]
        src_tokens = net_input["src_tokens"]
        dev = src_tokens.device
        bsz, src_len = src_tokens.shape[:2]
        beam, V = self.beam_size, self.vocab_size
        max_len = int(self.max_len_a * src_len + self.max_len_b)
        assert self.min_len <= max_len, "min_len cannot be larger than max_len, please adjust these!"
        # The encoder output is NOT replicated per beam (the reference does: sequence_generator.py:262-266): the decoder keeps
        # the cross-attention K / V once per sentence and maps beam rows onto them (ofa.py incremental path)
        enc = model.encoder.forward_torchscript(net_input)
        scores = torch.zeros(bsz * beam, max_len + 1, device=dev)
        static = None
        if self.cuda_graphs and dev.type == "cuda":
            # captured decoder steps hold weights-derived tensors (the concatenated q|k|v projection of qkv_eval): a weights
            # fingerprint (sum of the parameters' version counters: optimizers and load_state_dict bump them) retires the graphs
            # of older weights
            wv = sum(p._version for p in model.decoder.parameters())
            if getattr(self, "_weights_version", wv) != wv:
                self._static.clear()
            self._weights_version = wv
            sig = (bsz, beam, max_len, enc["encoder_out"][0].shape[0], str(enc["encoder_out"][0].dtype))
            static = self._static.get(sig)
            if static is None:
                static = self._static[sig] = {
                    # two token buffers: a step's bookkeeping gathers the surviving hypotheses from one into the other
                    "tokens": [torch.empty((bsz * beam, max_len + 2), dtype=torch.long, device=dev) for _ in range(2)],
                    "scores": [torch.zeros(bsz * beam, max_len + 1, device=dev) for _ in range(2)],
                    "ignore": [torch.zeros(bsz, beam, dtype=torch.bool, device=dev) for _ in range(2)],
                    "active": [torch.empty(bsz * beam, dtype=torch.long, device=dev) for _ in range(2)],
                    "eos_n": torch.zeros(bsz, dtype=torch.int32, device=dev),
                    "order": torch.zeros(bsz * beam, dtype=torch.long, device=dev),
                    "inc": {"_ofa_b200": {"reuse": True}}, "graphs": {}, "calls": 0, "pool": None}
            static["calls"] += 1
        if static is not None:
            tok_bufs = static["tokens"]
            for t_ in tok_bufs:
                t_.fill_(self.pad)
        else:
            tok_bufs = [torch.full((bsz * beam, max_len + 2), self.pad, dtype=torch.long, device=dev) for _ in range(2)]
        for t_ in tok_bufs:
            t_[:, 0] = self.bos
        tokens, alt_tokens = tok_bufs
        bsz0 = bsz
        fused = dev.type == "cuda"                      # one-launch bookkeeping (csrc/beam.cu beam_advance_kernel)
        if static is not None:                          # persistent buffers: captured steps hold pointers into them
            scores, alt_scores = static["scores"]
            cands_to_ignore, alt_ignore = static["ignore"]
            for t_ in (scores, alt_scores, cands_to_ignore, alt_ignore):
                t_.zero_()
            active_buf, eos_n = static["active"], static["eos_n"]
        else:
            alt_scores = torch.zeros_like(scores)
            cands_to_ignore = torch.zeros(bsz, beam, dtype=torch.bool, device=dev)
            alt_ignore = torch.zeros_like(cands_to_ignore)
            active_buf = [torch.empty(bsz * beam, dtype=torch.long, device=dev) for _ in range(2)]
            eos_n = torch.zeros(bsz, dtype=torch.int32, device=dev)
        if fused:
            if getattr(self, "_eos_host", None) is None or self._eos_host.numel() < bsz:
                self._eos_host = torch.zeros(max(bsz, 64), dtype=torch.int32).pin_memory()      # (page-locking is slow: once)
            eos_host = self._eos_host
        finalized: List[List[Dict]] = [[] for _ in range(bsz)]
        finished = [False] * bsz
        num_remaining = bsz
        cand_size = 2 * beam
        bbsz_offsets = (torch.arange(bsz, device=dev) * beam).unsqueeze(1)
        cand_offsets = torch.arange(cand_size, device=dev)
        inc = static["inc"] if static is not None else {}
        reorder_state = batch_idxs = None
        topk_ws = None
        trie = node = None
        if self.constraint_trie is not None:
            # CSR trie on the device; every row starts at the node reached by bos (the reference walks [0] + tokens[1:] from the
            # root at every step: sequence_generator.py:861-866)
            if self._trie_csr is None or self._trie_csr[0].device != dev:
                self._trie_csr = flatten_trie(self.constraint_trie, dev)
            trie = self._trie_csr
            root = torch.zeros(bsz * beam, dtype=torch.int32, device=dev)
            node = ops.trie_advance(trie, root, None, torch.zeros(bsz * beam, dtype=torch.long, device=dev))
        plen = node0 = None
        if prefix_tokens is not None:
            prefix_tokens = prefix_tokens.to(dev)
            if node is not None and not self.zero_shot:
                # the pre-softmax trie mask skips each sentence's forced prefix and walks [bos] + tokens[prefix_len + 1:]
                # (:862-868): -2 = unconstrained while the row is inside its prefix, the bos node at its end
                plen = prefix_tokens.ne(self.pad).sum(1).repeat_interleave(beam).to(torch.int32)
                node0 = node
                node = torch.where(plen > 0, torch.full_like(node0, -2), node0)
        if self.constraint_start is not None and trie is not None:
            raise ValueError("constraint_trie and constraint_range are mutually exclusive (sequence_generator.py:858,871)")
        for step in range(max_len + 1):
            if reorder_state is not None and batch_idxs is not None:
                corr = batch_idxs - torch.arange(batch_idxs.numel(), device=dev)
                reorder_state.view(-1, beam).add_(corr.unsqueeze(-1) * beam)
            # forced prefix tokens of this step (possibly of different lengths: pad = none for that sentence; :372-380)
            ptoks = None
            if prefix_tokens is not None and step < prefix_tokens.size(1) and step < max_len:
                ptoks = prefix_tokens[:, step].unsqueeze(-1).repeat(1, beam).view(-1).contiguous()
            graphable = (static is not None and step >= 1 and bsz == bsz0 and batch_idxs is None and static["calls"] >= 2
                         and (tokens is static["tokens"][0] or tokens is static["tokens"][1]))
            # the whole step as one graph (decoder + fused tail + bookkeeping) when every buffer it touches is persistent
            # (a prefix step keeps its tail eager -- the tail may synchronise -- but its decoder pass is captured like any other:
            # a captured step reads the cache state its predecessor left in the graph pool)
            whole = graphable and fused and step < max_len and node is None and ptoks is None and scores is static["scores"][
                0 if tokens is static["tokens"][0] else 1]
            if whole and not (cands_to_ignore is static["ignore"][0] or cands_to_ignore is static["ignore"][1]):
                par = 0 if tokens is static["tokens"][0] else 1
                static["ignore"][par].copy_(cands_to_ignore)       # (left behind by a finalisation step)
                cands_to_ignore, alt_ignore = static["ignore"][par], static["ignore"][1 - par]

            def tail(logits, ws, tokens=tokens, scores=scores, step=step, ptoks=ptoks):
                """fused tail: temperature, constraints, log-softmax, masks, n-gram blocking, + beam scores, top 2*beam.
                `ws`: workspace of an earlier eager call or None -- a captured step must NOT hold a pointer to an eager tensor
                that dies with this generate call, so the graph path allocates its own inside the capture."""
                lg = logits[:, -1, :]
                pfill = None
                if ptoks is not None:
                    if bool(ptoks.eq(self.eos).any()):
                        # a prefix that ends here: every beam of that sentence continues from the first one (:614-634); the
                        # log-probs follow from the replicated logits row (same prefix token for every beam of a sentence)
                        eb = ptoks.eq(self.eos).view(-1, beam)[:, 0]
                        first = tokens.view(bsz, beam, -1)[eb][:, 0, 1:step + 1]
                        assert bool((first == prefix_tokens[eb][:, :step]).all())
                        for t_ in (tokens, scores, lg):
                            v_ = t_.view(bsz, beam, -1)
                            v_[eb] = v_[eb][:, :1, :].expand(-1, beam, -1).clone()
                    if trie is None:
                        # every other token of a prefixed row: min over ALL rows of the prefix log-probs - 1 (:607-608)
                        lt = lg if self.temperature == 1.0 else lg.float() / self.temperature
                        R_ = lt.shape[0]
                        plp = ops.trie_score(lt, torch.arange(R_ + 1, dtype=torch.int32, device=dev),
                                             torch.full((R_,), -1, dtype=torch.int32, device=dev), ptoks, None, -1)
                        pfill = (plp.min() - 1).reshape(1)
                prev = scores.view(bsz, beam, -1)[:, :, step - 1].reshape(-1).contiguous() if step > 0 else None
                return ops.beam_topk(
                    lg, beam, min(cand_size, beam * V - 1), self.temperature, prev, step0=(step == 0), eos=self.eos, pad=self.pad,
                    unk=self.unk, unk_penalty=self.unk_penalty, block_eos=step < self.min_len and ptoks is None,
                    force_eos=step >= max_len,
                    eos_one=self.ignore_eos,
                    crange=(self.constraint_start, self.constraint_end) if self.constraint_start is not None else None,
                    range_post=self.zero_shot, trie=trie, node=node, trie_post=self.zero_shot, tokens=tokens, step=step,
                    ngram=self.no_repeat_ngram_size, ws=ws, prefix_tok=ptoks, prefix_fill=pfill)

            def advance(cand_scores, idx, act):
                ops.beam_advance(cand_scores, idx, cands_to_ignore.contiguous(), tokens, scores, alt_tokens, alt_scores, alt_ignore,
                                 act, eos_n[:bsz], beam, V, self.eos, step)

            act = active_buf[step & 1][:bsz * beam] if fused else None
            advanced = False
            if whole:
                alt_tokens = static["tokens"][1 if tokens is static["tokens"][0] else 0]
                alt_scores = static["scores"][1 if scores is static["scores"][0] else 0]
                logits, cand_scores, idx = self._graphed_step(model, static, step, tokens, enc, inc, reorder_state,
                                                              tail=tail, advance=lambda cs_, ix_: advance(cs_, ix_, act))
                advanced = True
            else:
                if graphable:
                    logits = self._graphed_step(model, static, step, tokens, enc, inc, reorder_state)
                else:
                    if reorder_state is not None:
                        model.decoder.reorder_incremental_state_scripting(inc, reorder_state)
                    logits, _ = model.decoder(tokens[:, :step + 1], encoder_out=enc, incremental_state=inc, padded_logits=True)
                cand_scores, idx, topk_ws = tail(logits, topk_ws)
            if fused and step < max_len:
                # the common step -- no hypothesis ends: one launch gathers the surviving hypotheses into the other buffers and
                # counts the eos candidates; the count is the one host synchronisation of the step (the reference synchronises in
                # masked_select, :447-450).  Any eos candidate -> the reference's own sequence of operations below, on the
                # untouched inputs.
                if not advanced:
                    if alt_tokens.shape != tokens.shape:
                        alt_tokens = torch.full_like(tokens, self.pad)
                        alt_scores, alt_ignore = torch.zeros_like(scores), torch.zeros_like(cands_to_ignore)
                    advance(cand_scores, idx, act)
                eos_host[:bsz].copy_(eos_n[:bsz], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                if int(eos_host[:bsz].sum()) == 0:
                    tokens, alt_tokens = alt_tokens, tokens
                    scores, alt_scores = alt_scores, scores
                    cands_to_ignore, alt_ignore = alt_ignore, cands_to_ignore
                    if node is not None:
                        node = self._next_nodes(trie, node, act, tokens[:, step + 1], plen, node0, step + 1)
                    reorder_state, batch_idxs = act, None
                    continue
            cand_beams = idx // V
            cand_indices = idx.fmod(V)
            cand_bbsz_idx = cand_beams + bbsz_offsets
            eos_mask = cand_indices.eq(self.eos) & cand_scores.ne(-math.inf)
            eos_mask[:, :beam][cands_to_ignore] = False
            eos_bbsz_idx = torch.masked_select(cand_bbsz_idx[:, :beam], eos_mask[:, :beam])
            finalized_sents: List[int] = []
            if eos_bbsz_idx.numel() > 0:
                eos_scores = torch.masked_select(cand_scores[:, :beam], eos_mask[:, :beam])
                finalized_sents = self._finalize(step, eos_bbsz_idx, eos_scores, tokens, scores, finalized, finished,
                                                 beam, max_len)
                num_remaining -= len(finalized_sents)
            if num_remaining == 0:
                break
            assert step < max_len, f"{step} < {max_len}"
            if len(finalized_sents) > 0:
                new_bsz = bsz - len(finalized_sents)
                batch_mask = torch.ones(bsz, dtype=torch.bool, device=dev)
                batch_mask[finalized_sents] = False
                batch_idxs = torch.arange(bsz, device=dev).masked_select(batch_mask)
                eos_mask = eos_mask[batch_idxs]
                cand_beams = cand_beams[batch_idxs]
                bbsz_offsets = bbsz_offsets[:new_bsz]
                cand_bbsz_idx = cand_beams + bbsz_offsets
                cand_scores = cand_scores[batch_idxs]
                cand_indices = cand_indices[batch_idxs]
                cands_to_ignore = cands_to_ignore[batch_idxs]
                if node is not None:
                    node = node.view(bsz, beam)[batch_idxs].reshape(-1).contiguous()
                if plen is not None:
                    plen = plen.view(bsz, beam)[batch_idxs].reshape(-1).contiguous()
                    node0 = node0[:new_bsz * beam]
                if prefix_tokens is not None:
                    prefix_tokens = prefix_tokens[batch_idxs]                     # :507-508
                scores = scores.view(bsz, -1)[batch_idxs].view(new_bsz * beam, -1)
                tokens = tokens.view(bsz, -1)[batch_idxs].view(new_bsz * beam, -1)
                bsz = new_bsz
            else:
                batch_idxs = None
            eos_mask[:, :beam] = ~((~cands_to_ignore) & (~eos_mask[:, :beam]))
            active_mask = eos_mask.long() * cand_size + cand_offsets[: eos_mask.size(1)]
            new_ignore, active_hypos = torch.topk(active_mask, k=beam, dim=1, largest=False)
            cands_to_ignore = new_ignore.ge(cand_size)[:, :beam]
            active_bbsz_idx = torch.gather(cand_bbsz_idx, 1, active_hypos).view(-1)
            tokens[:, :step + 1] = torch.index_select(tokens[:, :step + 1], 0, active_bbsz_idx)
            tokens.view(bsz, beam, -1)[:, :, step + 1] = torch.gather(cand_indices, 1, active_hypos)
            if step > 0:
                scores[:, :step] = torch.index_select(scores[:, :step], 0, active_bbsz_idx)
            scores.view(bsz, beam, -1)[:, :, step] = torch.gather(cand_scores, 1, active_hypos)
            if node is not None:       # trie node of every surviving hypothesis: its parent's node advanced by the chosen token
                node = self._next_nodes(trie, node, active_bbsz_idx.contiguous(), tokens[:, step + 1], plen, node0, step + 1)
            reorder_state = active_bbsz_idx
        for s in range(len(finalized)):
            # (:589-597 reads every score back with .item(): one device synchronisation per hypothesis; the host copies taken
            # once per finalisation step in _finalize give the same order)
            sc = torch.tensor([h.pop("_score_host") for h in finalized[s]])
            _, o = torch.sort(sc, descending=True)
            finalized[s] = [finalized[s][i] for i in o]
        return finalized

    def _graphed_step(self, model, static, step, tokens, enc, inc, reorder_state, tail=None, advance=None):
        """Beam reorder + one decoder step (+ with `tail` / `advance` the fused beam tail and the bookkeeping launch) as a CUDA
        graph.  The graph holds pointers into the persistent token / score / flag buffers, the reorder-index buffer and the
        decoder's KV-cache state; the Python side of that state (ping-pong index, length, the group -> sentence map tensor) is
        restored to its post-step value after every replay.  Returns logits, or (logits, cand_scores, cand_index) with a tail."""
        st = inc["_ofa_b200"]
        static["order"].copy_(reorder_state)
        key = (step, st.get("layout", 0), st["tcur"], st["ppar"], 0 if tokens is static["tokens"][0] else 1, tail is not None)
        g = static["graphs"].get(key)
        if g is None:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            cand = None
            with torch.cuda.graph(graph, pool=static["pool"]):
                model.decoder.reorder_incremental_state_scripting(inc, static["order"])
                logits, _ = model.decoder(tokens[:, :step + 1], encoder_out=enc, incremental_state=inc, padded_logits=True)
                if tail is not None:
                    cs_, ix_, _ws = tail(logits, None)
                    advance(cs_, ix_)
                    cand = (cs_, ix_)
            if static["pool"] is None:
                static["pool"] = graph.pool()
            g = static["graphs"][key] = {"graph": graph, "logits": logits, "cand": cand,
                                         "post": {k: st[k] for k in self._STATE_KEYS}}
        g["graph"].replay()
        st.update(g["post"])
        if tail is not None:
            return g["logits"], g["cand"][0], g["cand"][1]
        return g["logits"]

    @staticmethod
    def _next_nodes(trie, node, parent, tok, plen, node0, n_tokens):
        """Trie node of every surviving hypothesis: its parent's node advanced by the chosen token; with forced prefixes the
        walk starts at the end of the sentence's prefix (n_tokens = generated tokens so far, bos excluded; :862-868)."""
        nxt = ops.trie_advance(trie, node, parent, tok)
        if plen is not None:
            nxt = torch.where(plen > n_tokens, torch.full_like(nxt, -2), torch.where(plen == n_tokens, node0, nxt))
        return nxt

    _STATE_KEYS = ("tcur", "ppar", "len", "sent_row", "rows", "reordered_at")

    def _finalize(self, step, bbsz_idx, eos_scores, tokens, scores, finalized, finished, beam, max_len):
        tokens_clone = tokens.index_select(0, bbsz_idx)[:, 1:step + 2].clone()
        tokens_clone[:, step] = self.eos
        pos_scores = scores.index_select(0, bbsz_idx)[:, :step + 1].clone()
        pos_scores[:, step] = eos_scores
        pos_scores[:, 1:] = pos_scores[:, 1:] - pos_scores[:, :-1]
        if self.normalize_scores:
            eos_scores = eos_scores / (step + 1) ** self.len_penalty
        cum_unfin, prev = [], 0
        for f in finished:
            if f:
                prev += 1
            else:
                cum_unfin.append(prev)
        cum = torch.tensor(cum_unfin, dtype=torch.long, device=bbsz_idx.device)
        unfin_idx = bbsz_idx // beam
        sent = unfin_idx + cum.index_select(0, unfin_idx)
        sent_l, unfin_l = sent.tolist(), unfin_idx.tolist()
        score_l = eos_scores.tolist()
        toks, scs, poss = tokens_clone.unbind(0), eos_scores.unbind(0), pos_scores.unbind(0)     # one call each, not one per row
        empty = torch.empty(0)
        for i in range(bbsz_idx.numel()):
            if len(finalized[sent_l[i]]) < beam:
                finalized[sent_l[i]].append({"tokens": toks[i], "score": scs[i], "_score_host": score_l[i],
                                             "attention": empty, "alignment": empty, "positional_scores": poss[i]})
        newly = []
        for s, u in sorted(set(zip(sent_l, unfin_l))):
            if not finished[s] and (len(finalized[s]) == beam or step == max_len):
                finished[s] = True
                newly.append(u)
        return newly

"""The CPU oracle (oracle/ofa_oracle.py) against the fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  Tolerances: the oracle executes the same ATen ops as the reference, so fp32
agreement is to rounding (1e-5 relative on loss / grad norms, 1e-5 abs on logits)."""
import copy
import random

import pytest
import torch

from oracle import ofa_oracle as oo, synth
from tests.helpers import load_golden, build_case, tie, ZERO_GRAD_SUFFIXES

TRAIN_CASES = ["micro_pad", "micro_text_only", "micro_nomask_row", "micro_constraint", "micro_rdrop_sample",
               "micro_multitask_rdrop", "micro_plainflags", "micro_frozenbn_eval", "c1_tiny"]


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_train_case(name):
    fx = load_golden(name)
    case = fx["case"]
    cfg, sd, samples = build_case(case)
    ck = case["crit"]
    sdo = tie(sd)
    if "py_seed" in case:
        random.seed(case["py_seed"])
    inp = copy.deepcopy(samples)
    if case.get("sample_patch_num"):
        inp[0]["net_input"]["sample_patch_num"] = case["sample_patch_num"]
    stats = {}
    loss, ss, lg = oo.criterion_forward(
        sdo, cfg, inp if len(inp) > 1 else inp[0], epsilon=ck["label_smoothing"],
        use_rdrop=ck.get("use_rdrop", False), reg_alpha=ck.get("reg_alpha", 1.0),
        sample_patch_num=ck.get("sample_patch_num", 0), training=not case.get("eval_mode", False),
        stats_out=stats)
    (loss / ss).backward()
    assert ss == fx["sample_size"]
    assert abs(float(loss.detach()) - fx["loss"]) <= 1e-5 * abs(fx["loss"])
    if "logits_sub" in fx:
        got = lg["logits"].detach()
        assert (got[:, :, ::37] - fx["logits_sub"]).abs().max() < 1e-5
        assert (torch.logsumexp(got, -1) - fx["logits_lse"]).abs().max() < 1e-5
        po = fx["patch_orders"]
        if po is not None:
            assert torch.equal(po, lg["patch_orders"])
    else:
        for a, b in zip(lg["tasks"], fx["task_loss"]):
            assert abs(float(a["loss"]) - b) <= 1e-5 * abs(b)
    tot = 0.0
    for n, g in fx["grad_norms"].items():
        go = sdo[n].grad
        if g is None:
            assert go is None or float(go.norm()) == 0.0, n
            continue
        tot += float(go.norm()) ** 2
        if n.endswith(ZERO_GRAD_SUFFIXES):
            continue
        assert abs(float(go.norm()) - g) <= 1e-4 * g + 1e-9, n
    assert abs(tot ** 0.5 - fx["grad_norm_total"]) <= 1e-5 * fx["grad_norm_total"]
    if "bn_running_norms" in fx and stats and len(samples) == 1 and not ck.get("use_rdrop"):
        for k, v in stats.items():
            assert abs(float(v.norm()) - fx["bn_running_norms"][k]) <= 1e-4 * fx["bn_running_norms"][k], k


@pytest.mark.parametrize("name", ["gen_micro", "gen_micro_ngram", "gen_tiny", "gen_micro_trie", "gen_micro_trie_zeroshot",
                                  "gen_micro_range", "gen_micro_range_zeroshot", "gen_micro_prefix_trie", "gen_micro_prefix"])
def test_beam_search(name):
    fx = load_golden(name)
    case = fx["case"]
    cfg = synth.make_cfg(case["arch"], **case["cfg"])
    sd = synth.synth_state_dict(cfg, seed=0, **{k: case[k] for k in ("emb_std", "w_std") if k in case})
    sample = synth.make_batch(**case["batch"])
    g = case["gen"]
    trie = None
    if "trie" in case:
        trie = oo.Trie(2)
        for w in synth.trie_words(vocab=cfg.vocab_size, **case["trie"]):
            trie.insert(w)
    hyp = oo.generate(sd, cfg, sample["net_input"], beam=g["beam_size"], max_len_a=g["max_len_a"],
                      max_len_b=g["max_len_b"], min_len=g["min_len"],
                      no_repeat_ngram_size=g.get("no_repeat_ngram_size", 0), temperature=g.get("temperature", 1.0),
                      unk_penalty=g.get("unk_penalty", 0.0), constraint_trie=trie,
                      constraint_range=g.get("constraint_range"), zero_shot=g.get("zero_shot", False),
                      prefix_tokens=synth.prefix_tokens(vocab=cfg.vocab_size, **case["prefix"]) if "prefix" in case else None)
    assert len(hyp) == len(fx["tokens"])
    for s in range(len(hyp)):
        assert len(hyp[s]) == len(fx["tokens"][s])
        for h, t, sc in zip(hyp[s], fx["tokens"][s], fx["scores"][s]):
            assert torch.equal(h["tokens"], t)          # bit-exact token ids
            assert abs(float(h["score"]) - sc) < 1e-5


@pytest.mark.parametrize("name", ["allcand_micro", "allcand_tiny"])
def test_all_candidate_scores(name):
    """Oracle restatement of utils/eval_utils.py:161-214 against the scores of the reference's own eval_vqa_gen (golden)."""
    from oracle.make_golden import allcand_prompts
    fx = load_golden(name)
    case = fx["case"]
    cfg = synth.make_cfg(case["arch"], **case["cfg"])
    sd = synth.synth_state_dict(cfg, seed=0, **{k: case[k] for k in ("emb_std", "w_std") if k in case})
    sample = synth.make_batch(**case["batch"])
    answers = synth.candidate_answers(vocab=cfg.vocab_size, **case["answers"])
    trie = oo.Trie(2)
    for a in answers:
        trie.insert([0] + a.tolist() + [2])
    sc = oo.score_all_candidates(sd, cfg, sample["net_input"], allcand_prompts(case, cfg.vocab_size), answers, trie,
                                 case["valid_batch_size"])
    assert (sc - fx["scores"]).abs().max().item() < 1e-4
    assert sc.argmax(1).tolist() == fx["predicts"]


def test_bucket_tables_bit_exact():
    """Integer tables must be bit-exact (north_star); closed forms vs the reference builders' output."""
    import json, os
    from tests.helpers import GOLDEN
    spec = json.load(open(os.path.join(GOLDEN, "state_dict_spec.json")))
    for arch in spec:
        cfg = synth.make_cfg(arch)
        mine = synth.state_spec(cfg)
        assert [e[0] for e in spec[arch]["entries"]] == list(mine.keys())
        for (k, shape, dt) in spec[arch]["entries"]:
            assert tuple(shape) == tuple(mine[k][0]), k
    t = synth.token_bucket_table(256)
    assert t.shape == (1024, 1024) and int(t.min()) == 0 and int(t.max()) == 510
    assert int(t[0, 0]) == 255 and int(t[5, 0]) == 260 and int(t[0, 5]) == 250
    b = synth.image_bucket_table(42, 83 * 83 + 3)
    assert int(b[1, 1]) == 3444 and int(b[1, 2]) == 3443 and int(b[1, 43]) == 3361   # SURVEY.md 7 hard part 2
    assert int(b[0, 5]) == 6889 and int(b[5, 0]) == 6890 and int(b[0, 0]) == 6891

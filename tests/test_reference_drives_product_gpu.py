"""The UNMODIFIED reference criterion and the UNMODIFIED reference SequenceGenerator driving musketeer_b200.OFAModel on the
GPU: the product is a drop-in for the callers on either side of the hot path
(criterions/label_smoothed_cross_entropy.py:204-226, models/sequence_generator.py:209-598,776-907).

Needs the reference tree (baseline/_ref -- placed there by tools/install_reference.py / __graft_entry__.build() -- or
/root/reference) and runs in a fresh interpreter with the fairseq stand-ins on the path, so that the product binds to the same
`fairseq` base classes the reference's generator checks with isinstance."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_dir():
    for c in (os.environ.get("MUSKETEER_REF"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if c and os.path.isdir(os.path.join(c, "models", "ofa")):
            return c
    return None


def _run(body):
    ref = _ref_dir()
    if ref is None:
        pytest.skip("reference tree not present (baseline/_ref)")
    code = "import sys\nsys.path.insert(0, %r)\n" % ROOT + textwrap.dedent(body)
    env = dict(os.environ, MUSKETEER_REF=ref)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    return r.stdout


def test_reference_criterion_drives_product_model():
    """fp32: loss (1e-3 rel... observed ~1e-6), sample size and total gradient norm of reference-criterion(product-model) equal
    the golden values of reference-criterion(reference-model), incl. a constraint-masked task (in-place masking of the logits
    the model returned, label_smoothed_cross_entropy.py:233-236) and R-Drop."""
    out = _run("""
        import copy, random, torch
        from oracle import ref_harness as rh
        rh.load()                                   # reference modules + fairseq stand-ins on sys.path
        from tests.helpers import load_golden, build_case, build_product, to_device
        import musketeer_b200.plugin                # registers against the stand-in fairseq
        for name in ("micro_pad", "micro_constraint", "micro_rdrop_sample"):
            fx = load_golden(name)
            case = fx["case"]
            cfg, sd, samples = build_case(case)
            model, task = build_product(cfg, sd, dtype=torch.float32)
            model.train()
            crit = rh.build_criterion(task, **case["crit"])
            smp = to_device(copy.deepcopy(samples[0]), "cuda")
            if case.get("sample_patch_num"):
                smp["net_input"]["sample_patch_num"] = case["sample_patch_num"]
                po = fx["patch_orders"]
                model.encoder.patch_orders_override = po[0] if isinstance(po, list) else po
            loss, ss, log = crit(model, smp)
            (loss / ss).backward()
            gn = sum(float(p.grad.norm()) ** 2 for p in model.parameters() if p.grad is not None) ** 0.5
            rl = abs(float(loss) - fx["loss"]) / abs(fx["loss"])
            rg = abs(gn - fx["grad_norm_total"]) / fx["grad_norm_total"]
            print(name, "loss rel %.2e grad-norm rel %.2e" % (rl, rg))
            assert ss == fx["sample_size"] and rl < 1e-3 and rg < 1e-3, (name, rl, rg)
        print("ok")
    """)
    assert out.strip().endswith("ok"), out


def test_reference_generator_drives_product_model():
    """The reference SequenceGenerator (EnsembleModel, incremental decoding gated on FairseqIncrementalDecoder, encoder output
    replicated per beam and reordered every step, reorder_incremental_state_scripting) on the product model: tokens bit-exact
    against the golden hypotheses of reference-generator(reference-model)."""
    out = _run("""
        import copy, torch
        from oracle import ref_harness as rh, synth
        ns = rh.load()
        from tests.helpers import load_golden, build_product, to_device
        import musketeer_b200.plugin
        from fairseq.models import FairseqIncrementalDecoder
        for name in ("gen_micro_varied", "gen_micro_varied_ngram"):
            fx = load_golden(name)
            case = fx["case"]
            cfg = synth.make_cfg(case["arch"], **case["cfg"])
            sd = synth.synth_state_dict(cfg, seed=0, **{k: case[k] for k in ("emb_std", "w_std") if k in case})
            model, task = build_product(cfg, sd, dtype=torch.float32)
            model.eval()
            assert isinstance(model.decoder, FairseqIncrementalDecoder)
            gen = ns.sg.SequenceGenerator([model], task.target_dictionary, **case["gen"])
            assert gen.model.has_incremental_states()
            hyp = gen.generate([model], to_device(synth.make_batch(**case["batch"]), "cuda"))
            for s in range(len(hyp)):
                assert len(hyp[s]) == len(fx["tokens"][s])
                for h, t, sc in zip(hyp[s], fx["tokens"][s], fx["scores"][s]):
                    assert torch.equal(h["tokens"].cpu(), t), (name, s, h["tokens"].tolist(), t.tolist())
                    assert abs(float(h["score"]) - sc) < 1e-4
            print(name, "tokens equal")
        print("ok")
    """)
    assert out.strip().endswith("ok"), out

"""Synthetic Musketeer batches in the reference's `sample` layout (data/mm_data/*_dataset.py collaters), used by
bench.py and smoke runs.  Shapes follow SURVEY.md 8(d) C2: TEP prompts measured with the GPT-2 BPE."""
import torch

PAD, BOS, EOS = 1, 0, 2
VOCAB = 59457

# task -> (src_len, tgt_len, has_image, target_prefix_pad)   (caption / VQA / VG / SNLI-VE / gigaword)
TEP_TASKS = [
    ("caption", 137, 12, True, 0),
    ("vqa", 230, 232, True, 171),       # prompt copied into the decoder (prompt_type=prev_output), target padded over it
    ("refcoco", 259, 5, True, 0),
    ("snli_ve", 250, 250, True, 217),
    ("gigaword", 185, 12, False, 0),
]


def make_task_batch(bsz, src_len, tgt_len, img, has_image, prefix_pad, seed, vocab=VOCAB, pin=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    hi = min(50265, vocab)
    src = torch.randint(4, hi, (bsz, src_len), generator=g)
    prev = torch.randint(4, hi, (bsz, tgt_len), generator=g)
    tgt = torch.randint(4, hi, (bsz, tgt_len), generator=g)
    src[:, 0] = BOS
    prev[:, 0] = BOS
    src[:, -1] = EOS
    tgt[:, -1] = EOS
    for b in range(bsz):          # ragged right padding like a real collated batch
        p = min(3 * (b % 4), src_len - 3)
        if p:
            src[b, src_len - p:] = PAD
            src[b, src_len - p - 1] = EOS
    if prefix_pad:
        tgt[:, :prefix_pad] = PAD
    sample = {"nsentences": bsz, "ntokens": int(tgt.ne(PAD).sum()), "target": tgt,
              "net_input": {"src_tokens": src, "src_lengths": src.ne(PAD).sum(1), "prev_output_tokens": prev}}
    if has_image:
        sample["net_input"]["patch_images"] = torch.randn(bsz, 3, img, img, generator=g)
        sample["net_input"]["patch_masks"] = torch.ones(bsz, dtype=torch.bool)
    if pin:
        sample = map_tensors(sample, lambda t: t.pin_memory())
    return sample


def make_tep_group(task_batch, img=384, seed=0, vocab=VOCAB, pin=False):
    """One Musketeer micro-step: a list with one batch per task (tasks/mm_tasks/musketeer_task.py:517-541)."""
    return [make_task_batch(task_batch, s, t, img, im, pp, seed * 100 + i, vocab, pin)
            for i, (_, s, t, im, pp) in enumerate(TEP_TASKS)]


def map_tensors(obj, fn):
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: map_tensors(v, fn) for k, v in obj.items()}
    if isinstance(obj, list):
        return [map_tensors(v, fn) for v in obj]
    return obj


def to_device(sample, device, float_dtype):
    return map_tensors(sample, lambda t: t.to(device=device, dtype=float_dtype, non_blocking=True)
                       if t.is_floating_point() else t.to(device, non_blocking=True))


def batch_bytes(sample, float_dtype=None):
    n = [0]

    def acc(t):
        n[0] += t.numel() * t.element_size()
        return t
    map_tensors(sample, acc)
    return n[0]


class Dictionary:
    """Minimal stand-in for the fairseq dictionary (len 59457; pad 1, bos 0, eos 2, unk 3: tasks/ofa_task.py:93-116)."""
    def __init__(self, n=VOCAB): self.n = n
    def __len__(self): return self.n
    def pad(self): return 1
    def eos(self): return 2
    def bos(self): return 0
    def unk(self): return 3
    def __eq__(self, o): return isinstance(o, Dictionary) and o.n == self.n
    def __ne__(self, o): return not self.__eq__(o)


class Task:
    def __init__(self, n=VOCAB):
        d = Dictionary(n)
        self.source_dictionary = self.target_dictionary = self.tgt_dict = self.src_dict = d


def build_model(arch="ofa_base", device="cuda", dtype=torch.bfloat16, seed=0, vocab=VOCAB, **over):
    """Random-init OFA with the Musketeer flag set (run_scripts/musketeer/train_musketeer.sh:124-176), dropout 0."""
    from types import SimpleNamespace
    from . import ARCHS, OFAModel
    kw = dict(scale_attn=True, scale_fc=True, scale_heads=True, add_type_embedding=True,
              disable_entangle=True, layernorm_embedding=True, patch_layernorm_embedding=True,
              code_layernorm_embedding=True, share_all_embeddings=True, encoder_normalize_before=True,
              decoder_normalize_before=True, dropout=0.0, attention_dropout=0.0, patch_image_size=384)
    kw.update(over)             # e.g. dropout=0.1, encoder_drop_path_rate=0.1, decoder_drop_path_rate=0.1 (the script's values)
    args = SimpleNamespace(**kw)
    ARCHS[arch](args)
    torch.manual_seed(seed)
    task = Task(vocab)
    model = OFAModel.build_model(args, task)
    return model.to(device=device, dtype=dtype), task

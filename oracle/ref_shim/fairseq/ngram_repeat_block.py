"""fairseq.ngram_repeat_block.NGramRepeatBlock — pure-python path of upstream (no CUDA ext)."""
import torch
from torch import nn

class NGramRepeatBlock(nn.Module):
    def __init__(self, no_repeat_ngram_size, use_extension=False):
        super().__init__()
        self.no_repeat_ngram_size = no_repeat_ngram_size

    def forward(self, tokens, lprobs, bsz, beam_size, step):
        n = self.no_repeat_ngram_size
        gen_ngrams = [dict() for _ in range(bsz * beam_size)]
        cpu_tokens = tokens.cpu()
        for bbsz_idx in range(bsz * beam_size):
            gen_tokens = cpu_tokens[bbsz_idx].tolist()
            for ngram in zip(*[gen_tokens[i:] for i in range(n)]):
                key = ",".join(str(x) for x in ngram[:-1])
                gen_ngrams[bbsz_idx][key] = gen_ngrams[bbsz_idx].get(key, []) + [ngram[-1]]
        if step + 2 - n >= 0:
            banned = []
            for bbsz_idx in range(bsz * beam_size):
                key = ",".join(str(x) for x in cpu_tokens[bbsz_idx, step + 2 - n: step + 1].tolist())
                banned.append(gen_ngrams[bbsz_idx].get(key, []))
        else:
            banned = [[] for _ in range(bsz * beam_size)]
        for bbsz_idx in range(bsz * beam_size):
            lprobs[bbsz_idx][torch.tensor(banned[bbsz_idx], dtype=torch.int64)] = torch.tensor(-float("inf")).to(lprobs)
        return lprobs

from torch.nn.modules.loss import _Loss
CRITERION_REGISTRY = {}
class FairseqCriterion(_Loss):
    def __init__(self, task):
        super().__init__()
        self.task = task
        if hasattr(task, "target_dictionary"):
            tgt_dict = task.target_dictionary
            self.padding_idx = tgt_dict.pad() if tgt_dict is not None else -100
def register_criterion(name, dataclass=None):
    def deco(cls):
        CRITERION_REGISTRY[name] = cls
        return cls
    return deco

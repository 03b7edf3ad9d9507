def fsdp_wrap(module, **kwargs):
    return module

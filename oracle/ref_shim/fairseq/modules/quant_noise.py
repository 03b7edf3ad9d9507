def quant_noise(module, p, block_size):
    assert p <= 0, "quant-noise is off in every Musketeer script"
    return module

import torch


def rel_l2(a, b):
    a = a.float().flatten()
    b = b.float().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def bf(x):
    return x.to(torch.bfloat16)


def rb(x):
    """round fp32 through bf16 (the reference rounds every op's output to bf16)"""
    return x.to(torch.bfloat16).float()

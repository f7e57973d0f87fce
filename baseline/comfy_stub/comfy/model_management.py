import torch


def soft_empty_cache(*_a, **_k):
    return None


def unload_all_models(*_a, **_k):
    return None


def get_torch_device():
    return torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")

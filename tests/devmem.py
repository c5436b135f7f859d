"""Where the device-pointer tests keep their buffers: CUDA memory through torch on the GPU box; plain host
memory when the suite runs against the host emulation of the kernels (ADSB_EMU=1, tests/emu)."""
import os

EMU = os.environ.get("ADSB_EMU") == "1"
DEVICE = "cpu" if EMU else "cuda"


def to_dev(t):
    return t.contiguous() if EMU else t.cuda()


def stream():
    if EMU:
        return 0
    import torch
    return torch.cuda.current_stream().cuda_stream


def sync():
    if not EMU:
        import torch
        torch.cuda.synchronize()

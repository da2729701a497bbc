"""Output stage of the reference CLI, restated (TEST INFRASTRUCTURE): scripts/evaluation/inference.py:153-159 /
funcs.py:208-216 -- clamp, map [-1, 1] to [0, 255], truncate to uint8, frames-major THWC for the video writer."""
import torch


def video_to_uint8(video):
    """video: [b, c, t, h, w] float -> uint8 [b, t, h, w, c] with the reference's torch ops in its order."""
    video = torch.clamp(video.detach().cpu().float(), -1., 1.)
    out = []
    for i in range(video.shape[0]):
        grid = video[i, ...]
        grid = (grid + 1.0) / 2.0
        out.append((grid * 255).to(torch.uint8).permute(1, 2, 3, 0))
    return torch.stack(out, 0)

"""Rate / distortion metrics of the codec (reference metrics.py:75-89 PSNR and BPP, :90-101 Metric), with
the per-iteration PSNR taken from the squared-error sums the fused tail kernel already reduces on the
device: one device->host copy for all T rates instead of T `.item()` syncs per logged batch
(train_model.py:128).  SSIM / MSSIM (metrics.py:9-72) are image-quality metrics outside the coding
loop and are not provided here."""
import numpy as np
import torch

from utils import recur


def PSNR(output, target, MAX=1.0):
    """metrics.py:75-81: 20 log10(MAX) - 10 log10(mse) over the whole batch"""
    with torch.no_grad():
        mse = torch.mean((output.float() - target.float()) ** 2)
        return (20 * torch.log10(torch.tensor(float(MAX), device=mse.device)) - 10 * torch.log10(mse)).item()


def BPP(code, img):
    """metrics.py:84-89: 8 * bytes of the packed code / pixels (all images, one channel)"""
    with torch.no_grad():
        return 8 * code.nbytes / (img.numel() / img.size(1))


def psnr_per_iteration(output, MAX=1.0):
    """[T] PSNR values from output['mse'] (device tensor written by the fused loop): identical to
    recur(PSNR, output['img'], input['img']) up to fp32 rounding of the reference's mean"""
    mse = output['mse'].detach().double().cpu().numpy()          # the one D2H copy
    return [float(20 * np.log10(MAX) - 10 * np.log10(m)) for m in mse]


def _psnr(input, output):
    if isinstance(output, dict) and 'mse' in output:
        return psnr_per_iteration(output)
    return recur(PSNR, output['img'], input['img'])


class Metric(object):
    def __init__(self):
        self.metric = {'Loss': (lambda input, output: output['loss'].item()),
                       'PSNR': _psnr,
                       'BPP': (lambda input, output: recur(BPP, output['code'], input['img']))}

    def evaluate(self, metric_names, input, output):
        evaluation = {}
        for metric_name in metric_names:
            if metric_name not in self.metric:
                raise ValueError('Not valid metric name (the B200 path provides Loss, PSNR, BPP)')
            evaluation[metric_name] = self.metric[metric_name](input, output)
        return evaluation

"""-m gpu, informational: the reference algorithm (the oracle's plain torch ops = cuDNN / cuBLAS eager kernels, fp32 with
PyTorch's default TF32 convolutions) on the same B200, next to this repo's path, on the bench workload.  It is the
"kernel to beat" of SURVEY section 8d; the numbers are printed for DESIGN.md, only sanity is asserted."""
import pytest
import torch

from conftest import make_inputs
from oracle import wav2sleep_oracle as oracle
from wav2sleep_b200 import build_default

pytestmark = pytest.mark.gpu
CARDIO = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}


def _time(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def test_eager_torch_baseline_on_the_same_gpu(cuda_device):
    B, S = 16, 1200
    model = build_default(CARDIO, 4, seed=0).to(cuda_device).eval()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    x = {k: v.to(cuda_device) for k, v in make_inputs(CARDIO, B, S, seed=42).items()}
    cfg = oracle.cardio_config()

    def eager():
        with torch.no_grad():
            z = oracle.signal_encoders(x, sd, cfg)
            m = oracle.epoch_mixer(z, sd, cfg)
            s = oracle.sequence_mixer(m, sd, cfg)
            return (s @ sd["classifier.weight"].t() + sd["classifier.bias"]).argmax(-1)

    def ours():
        with torch.inference_mode():
            return model.predict(x)

    t_eager, t_ours = _time(eager), _time(ours, n=5)
    agree = (eager() == ours()).float().mean().item()
    print(f"eager torch (cuDNN/cuBLAS, fp32+TF32 convs) {t_eager:.1f} ms = {B * 10 / t_eager * 1e3:.0f} recording-h/s; "
          f"this repo {t_ours:.2f} ms = {B * 10 / t_ours * 1e3:.0f} recording-h/s; ratio {t_eager / t_ours:.1f}x; "
          f"argmax agreement between the two GPU paths {agree:.4f}")
    assert agree > 0.99 and t_ours < t_eager

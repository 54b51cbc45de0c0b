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


def test_eager_torch_training_step_baseline(cuda_device):
    """fwd + CE + bwd of the same graph through torch autograd (eager cuDNN / cuBLAS, fp32, no dropout) vs this repo's
    training path, 4 nights (the eager graph keeps every fp32 intermediate: 16 nights would not be a safe allocation)."""
    B, S = 4, 1200
    model = build_default(CARDIO, 4, seed=0).to(cuda_device)
    model.epoch_mixer.dropout = 0.0
    for blk in model.sequence_mixer.dilated_convs:
        blk.dropout.p = 0.0
    x = {k: v.to(cuda_device) for k, v in make_inputs(CARDIO, B, S, seed=42).items()}
    y = torch.randint(0, 4, (B, S), device=cuda_device)
    cfg = oracle.cardio_config()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}

    def eager():
        for p in params.values():
            p.grad = None
        z = oracle.signal_encoders(x, params, cfg)
        s = oracle.sequence_mixer(oracle.epoch_mixer(z, params, cfg), params, cfg)
        logits = s @ params["classifier.weight"].t() + params["classifier.bias"]
        torch.nn.functional.cross_entropy(logits.view(-1, 4), y.view(-1)).backward()

    model.train()

    def ours():
        model.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(model(x).view(-1, 4), y.view(-1)).backward()

    t_eager, t_ours = _time(eager, n=2), _time(ours, n=3)
    g_ref = params["signal_encoders.encoders.ECG.cnn.3.conv2.conv.weight"].grad
    g = model.signal_encoders.get_encoder("ECG").cnn[3].conv2.conv.weight.grad
    cos = torch.nn.functional.cosine_similarity(g.flatten().float(), g_ref.flatten(), dim=0).item()
    print(f"eager torch fwd+bwd {t_eager:.1f} ms vs this repo {t_ours:.1f} ms for {B} nights: {t_eager / t_ours:.1f}x; "
          f"peak eager memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB; cos(grad) {cos:.5f}")
    assert cos > 0.99 and t_ours < t_eager  # both sides are reduced precision here (TF32 convolutions vs fp16 storage)

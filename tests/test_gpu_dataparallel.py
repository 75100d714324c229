"""nn.DataParallel re-entrancy of the drop-in (train_distributed.py:72: the only live multi-GPU mechanism of the
reference): replicate -> scatter -> one worker thread per device -> gather, gradients reduced onto device 0.  Needs two
GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
import numpy as np
import pytest
import torch

from tests.util_gpu import build_model, load_pkg, make_case, oracle_step, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 CUDA devices")
@pytest.mark.parametrize("precision,ltol,gtol", [("fp32", 1e-5, 1e-5), ("bf16", 1e-3, 1e-2)])
def test_dataparallel_wrapped_dropin_matches_full_batch_oracle(precision, ltol, gtol):
    m = load_pkg()
    B, Z, H, L = 12, 16, 24, 2
    P, ids, onehot, eps = make_case(91, 92, B, Z, H, L)
    ref = oracle_step(P, onehot, eps, L, train=False)           # eval(): z = mu, so no per-replica normal draws to align
    model = build_model(m, P, Z, H, L, precision).eval()
    dp = torch.nn.DataParallel(model, device_ids=[0, 1])
    x = torch.from_numpy(onehot).cuda(0)
    m.models2d.max_len = 120
    for it in range(2):                                          # twice: replicas are rebuilt every forward, engines are not
        model.zero_grad()
        probs, mu, logvar = dp(x)                                # train_distributed.py:87
        assert probs.shape == (B, 120, 35) and probs.device.index == 0
        loss = m.loss_function(probs, x, mu, logvar)             # on the gathered full batch (:89)
        loss.backward()
        torch.cuda.synchronize()
        assert abs(float(loss) - ref["loss"]) <= ltol * abs(ref["loss"])
        bad = {k: rel_l2(p.grad.cpu().numpy(), ref["grads"][k]) for k, p in model.named_parameters()}
        bad = {k: e for k, e in bad.items() if not e <= gtol}
        assert not bad, (it, bad)
    assert sorted(model._engines) == ["cuda:0", "cuda:1"]        # one engine per device, built once
    np.testing.assert_allclose(probs.detach().cpu().numpy(), ref["probs"], rtol=2e-2 if precision == "bf16" else 2e-5, atol=1e-3 if precision == "bf16" else 1e-7)

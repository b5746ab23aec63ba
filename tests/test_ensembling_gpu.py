"""GPU: inference/ensembling.py wrappers (one-pass stack reduction, in-place activations) against the reference's semantics
restated with torch ops: output.mul(T).sigmoid_() / .softmax(dim) (ensembling.py:39-42,64-67) and
_deaugment_averaging(torch.stack(outputs), reduction) (ensembling.py:108-118, via oracle/tta_oracle.averaging)."""
import pytest
import torch
from torch import nn

from oracle import tta_oracle as AO
from pytorch_toolbelt_b200.inference.ensembling import (ApplySigmoidTo, ApplySoftmaxTo, Ensembler, PickModelOutput, SelectByIndex,
                                                        stack_reduce)

pytestmark = pytest.mark.gpu


class Toy(nn.Module):
    def __init__(self, seed, as_dict=True):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w = nn.Parameter(torch.randn(5, 3, 1, 1, generator=g))
        self.as_dict = as_dict

    def forward(self, x):
        y = torch.nn.functional.conv2d(x, self.w)
        return {"logits": y, "aux": y[:, :2] * 2} if self.as_dict else y


@pytest.mark.parametrize("reduction", ["mean", "sum", "gmean", "hmean", "harmonic1p", "logodd", "log1p"])
def test_ensembler_matches_stack_then_average(reduction):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 37, 41, generator=g).cuda()
    models = [Toy(s).cuda() for s in range(3)]
    positive = reduction not in ("mean", "sum")

    def prep(t):
        return torch.sigmoid(t) if positive else t

    wrapped = [nn.Sequential(m, SelectByIndex("logits")) for m in models]
    outs = [prep(w(x)) for w in wrapped]
    want = AO.averaging(torch.stack([o.detach().cpu() for o in outs]), reduction)
    got = stack_reduce(outs, reduction)
    torch.testing.assert_close(got.cpu(), want, rtol=2e-6, atol=1e-6)
    if reduction == "mean":
        ens = Ensembler(models, reduction="mean")
        res = ens(x)
        assert set(res) == {"logits", "aux"}
        torch.testing.assert_close(res["logits"].cpu(), torch.stack([m(x)["logits"].detach().cpu() for m in models]).mean(0), rtol=2e-6, atol=1e-6)
        only = Ensembler(models, reduction="mean", outputs=["aux"])(x)
        assert list(only) == ["aux"]
        plain = Ensembler([Toy(s, as_dict=False).cuda() for s in range(3)], reduction="sum")(x)
        assert torch.is_tensor(plain) and plain.shape == (2, 5, 37, 41)


def test_activation_wrappers_and_pickers():
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 19, 23, generator=g).cuda()
    model = Toy(7).cuda()
    raw = model(x)["logits"].detach()
    sig = ApplySigmoidTo(model, "logits", temperature=1.7)(x)["logits"]
    torch.testing.assert_close(sig, raw.mul(1.7).sigmoid(), rtol=2e-6, atol=1e-7)
    sm = ApplySoftmaxTo(model, ["logits"], dim=1, temperature=0.5)(x)["logits"]
    torch.testing.assert_close(sm, raw.mul(0.5).softmax(dim=1), rtol=2e-6, atol=1e-7)
    torch.testing.assert_close(model(x)["logits"], raw)  # the model's own output was not modified in place
    assert torch.equal(PickModelOutput(model, "aux")(x), model(x)["aux"])
    with pytest.raises(KeyError):
        stack_reduce([raw, raw], "median")
    with pytest.raises(RuntimeError):
        stack_reduce([raw, raw[:, :2]], "mean")

"""Generate tests/golden/variants.pt by running the UNMODIFIED reference modules from /root/reference
(container-only; the fixture is committed).  Run:  python tests/golden/make_golden_variants.py

Pins the model variants outside the default configuration:
  cbn        - tools_for_model.ComplexBatchNorm (tools_for_model.py:335-512): forward in train and eval mode,
               running-stat updates, and autograd gradients wrt the input and Wrr/Wri/Wii/Br/Bi;
  cbn_model  - DCCRN(use_cbn=True, use_clstm=True) (DCCRN.py:80-81): enhanced waveform in eval and train
               mode, -SI-SNR loss (tools_for_loss.py:37-47) and parameter gradients in train mode;
  lstm_model - DCCRN(use_clstm=False) (DCCRN.py:100-110, 193-199: 2-layer nn.LSTM + `tranform` Linear),
               same quantities.
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

KN = [4, 8, 8, 16, 16, 16]
GRAD_KEYS = {
    "cbn_model": ["encoder.0.1.Wri", "encoder.2.1.Wrr", "decoder.1.1.Bi", "decoder.4.1.Wii",
                  "encoder.1.0.real_conv.weight", "enhance.0.real_lstm.weight_ih_l0"],
    "lstm_model": ["enhance.weight_ih_l0", "enhance.weight_hh_l0", "enhance.bias_hh_l1", "enhance.weight_ih_l1",
                   "tranform.weight", "tranform.bias", "encoder.1.0.real_conv.weight", "decoder.0.0.imag_conv.weight"],
}


def model_case(mods, name, **kw):
    torch.manual_seed(21 if name == "cbn_model" else 33)
    m = mods["DCCRN"].DCCRN(masking_mode="E", kernel_num=KN, **kw)
    if name == "cbn_model":       # non-trivial running statistics for the eval-mode pin
        g0 = torch.Generator().manual_seed(77)
        for n_, b in m.named_buffers():
            if ".1.RV" in n_ and not n_.endswith("RVri"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g0))
            elif ".1.RM" in n_ or n_.endswith("RVri"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g0))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()
          if not k.startswith(("stft.", "istft."))}
    g = torch.Generator().manual_seed(6)
    x, y = 0.1 * torch.randn(3, 2400, generator=g), 0.1 * torch.randn(3, 2400, generator=g)
    out = {"kw": kw, "kernel_num": KN, "sd": sd, "x": x, "y": y}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.eval()
        with torch.no_grad():
            out["wav_eval"] = m(x)[-1].clone()
        m.train()
        wav = m(x)[-1]
        loss = -mods["tools_for_loss"].si_snr(wav, y)
        loss.backward()
    out["wav_train"] = wav.detach().clone()
    out["loss_train"] = float(loss)
    params = dict(m.named_parameters())
    out["grads"] = {k: params[k].grad.detach().clone() for k in GRAD_KEYS[name]}
    return out


def cbn_case(mods):
    g = torch.Generator().manual_seed(9)
    m = mods["tools_for_model"].ComplexBatchNorm(8)
    for n_, b in m.named_buffers():
        if b.is_floating_point():
            b.copy_(0.5 + torch.rand(b.shape, generator=g) if "RV" in n_ and "ri" not in n_
                    else 0.1 * torch.randn(b.shape, generator=g))
    m.Br.data.normal_(generator=g)
    m.Bi.data.normal_(generator=g)
    state = {k: v.detach().clone() for k, v in list(m.named_parameters()) + list(m.named_buffers())}
    x = torch.randn(3, 8, 4, 6, generator=g) + 0.2
    gy = torch.randn(3, 8, 4, 6, generator=g)
    out = {"state": state, "x": x, "gy": gy}
    for training in (True, False):
        m = mods["tools_for_model"].ComplexBatchNorm(8)      # fresh module: train mode ties the buffers to the graph
        m.load_state_dict(state)
        m.train(training)
        xa = x.clone().requires_grad_(True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            yv = m(xa)
            (yv * gy).sum().backward()
        key = "train" if training else "eval"
        out[key] = {"y": yv.detach().clone(), "dx": xa.grad.clone(),
                    "grads": {k: getattr(m, k).grad.clone() for k in ("Wrr", "Wri", "Wii", "Br", "Bi")},
                    "running": {k: getattr(m, k).detach().clone() for k in ("RMr", "RMi", "RVrr", "RVri", "RVii")}}
    return out


def main():
    torch.set_num_threads(4)
    mods = ref_shim.load()
    out = {"cbn": cbn_case(mods),
           "cbn_model": model_case(mods, "cbn_model", rnn_units=16, use_clstm=True, use_cbn=True),
           "lstm_model": model_case(mods, "lstm_model", rnn_units=24, use_clstm=False)}
    torch.save(out, os.path.join(HERE, "variants.pt"))
    print("variants.pt written:", {k: list(v) for k, v in out.items()})


if __name__ == "__main__":
    main()

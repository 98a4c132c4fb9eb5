"""TEST INFRASTRUCTURE — deterministic, implementation-independent weights for parity runs.

Both the reference Dreamer and pydreamer_b200.Dreamer expose the same state_dict keys/shapes; filling them from
a seeded generator (instead of each implementation's own init order) gives identical weights on both sides and
keeps the committed golden fixtures small (seeds + expected outputs, not tensors)."""
import torch


def seeded_state_dict(template, seed=0, dtype=torch.float32):
    """template: any state_dict with the reference key names; returns new tensors of the same shapes."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(template.keys()):
        shape = tuple(template[name].shape)
        r = torch.randn(shape, generator=g)
        if name.endswith("dummy"):
            v = torch.zeros(shape)
        elif len(shape) == 1:
            is_ln_weight = name.endswith("norm.weight") or (name.endswith(".weight"))
            v = 1.0 + 0.1 * r if is_ln_weight else 0.1 * r
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if "model.2.weight" in name or "model.4.weight" in name or "model.6.weight" in name or "model.8.weight" in name:
                if "decoder.image" in name:        # ConvTranspose2d (Cin, Cout, k, k): fan_in = Cin * k*k / stride^2
                    fan_in = shape[0] * shape[2] * shape[3] // 4
            v = r * (1.0 / max(fan_in, 1)) ** 0.5
        out[name] = v.to(dtype)
    return out

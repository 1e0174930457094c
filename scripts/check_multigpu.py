"""N-GPU parity check (run under torchrun): the image-sharded sweep + embed + gather must give, on every rank, exactly
the tensors of a single-process run (canonical top-k order makes values AND ids rank-count invariant).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/check_multigpu.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


class Images(torch.utils.data.Dataset):
    def __init__(self, n, kind):
        g = torch.Generator().manual_seed(5)
        self.u8 = torch.randint(0, 255, (n, 3, 224, 224), generator=g, dtype=torch.uint8)
        self.kind, self.name = kind, f"chk-{kind}-{n}"

    def __len__(self):
        return self.u8.shape[0]

    def __getitem__(self, i):
        return ((self.u8[i].float() / 255 - 0.45) / 0.23, 0) if self.kind == "model" else self.u8[i]


def main():
    import torchvision

    from semanticlens_b200.component_visualization import ActivationComponentVisualizer, aggregators
    from semanticlens_b200.foundation_models import OpenClip
    from semanticlens_b200.lens import Lens

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = torchvision.models.resnet18(weights=None).eval().to(dev)
    model.name = "rn18"
    n = 203  # not divisible by 2, 4 or 8
    ds_m, ds_f = Images(n, "model"), Images(n, "fm")
    fm = OpenClip("ViT-B-32", device=dev, load_weights=False, seed=1)
    layers = ["layer2", "layer4"]

    def run():
        cv = ActivationComponentVisualizer(model, ds_m, ds_f, layers, 7, device=dev, aggregate_fn=aggregators.aggregate_conv_mean)
        cv.show_progress = False
        cv.exchange = os.environ.get("SLB_EXCHANGE", "winners")
        db = Lens(fm, device=dev).compute_concept_db(cv, batch_size=32)
        return {k: (cv.actmax_cache.cache[k].activations.clone(), cv.actmax_cache.cache[k].sample_ids.clone(), db[k]) for k in layers}

    single = run()  # torch.distributed not initialised yet: the whole dataset on this GPU
    dist.init_process_group("nccl", device_id=dev)
    sharded = run()
    ok = True
    for k in layers:
        v1, i1, d1 = single[k]
        v2, i2, d2 = sharded[k]
        same_v = torch.equal(v1.view(torch.int16), v2.view(torch.int16))
        same_i = torch.equal(i1, i2)
        err = ((d1 - d2).abs().max() / d1.abs().max()).item()
        print(f"rank {dist.get_rank()}/{dist.get_world_size()} layer {k}: values {same_v} ids {same_i} concept_db rel diff {err:.2e}", flush=True)
        ok &= same_v and same_i and torch.equal(d1, d2)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if flag.item():
        raise SystemExit("MULTI-GPU PARITY FAILED")
    print("multi-gpu parity ok", flush=True)


if __name__ == "__main__":
    main()

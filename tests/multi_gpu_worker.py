"""torchrun worker for tests/test_gpu_multi.py: every rank post-processes its image shard on its own GPU, the packed
detections are gathered to rank 0 over NCCL, and rank 0 checks them bit for bit against its own full-batch run."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cerberusdet_b200 import ops  # noqa: E402
from cerberusdet_b200.pipeline import PostHeadPipeline  # noqa: E402
from cerberusdet_b200.shard import DetectionGatherer, GatherDelivery, PeerDelivery, gather_detections, make_delivery, shard_range  # noqa: E402
from cerberusdet_b200.synth import STRIDES, synth_heads  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ncs = [20, 19, 12]
    kw = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)

    def run(images):
        heads = synth_heads(images, ncs, 320, torch.float16, "iid", cfg=5)
        ys = ops.decode_heads([[x.to(dev) for x in lv] for lv in heads], STRIDES)
        return ops.nms_batched(ys, **kw)

    # (a) equal shards through the packed single-collective gatherer
    n_images = 8 * world
    mine = shard_range(n_images, rank, world)
    gat = DetectionGatherer(len(ncs), len(mine), kw["max_det"], dev, dst=0)
    heads = synth_heads(mine, ncs, 320, torch.float16, "iid", cfg=5)
    ys = ops.decode_heads([[x.to(dev) for x in lv] for lv in heads], STRIDES)
    ops.nms_batched(ys, out=gat.out, **kw)
    gat.launch()
    d, c = gat.result()
    # (b) ragged shards through gather_detections
    n2 = 8 * world + 3
    mine2 = shard_range(n2, rank, world)
    d2l, c2l = run(mine2)
    d2, c2 = gather_detections(d2l, c2l, dst=0, n_images=n2)
    # (c) the streaming engine with both per-batch deliveries: the NMS kernel writes straight into rank 0's peer-mapped
    # symmetric memory (no collective on the data path) / one asynchronous gather per batch; several batches per slot
    dev_heads = [[x.to(dev) for x in lv] for lv in heads]
    kinds = []
    streamed = {}
    for mode in ("peer", "peer_direct", "gather"):
        os.environ["CERB_DELIVERY"] = mode
        deliv = make_delivery(len(ncs), len(mine), kw["max_det"], dev)
        kinds.append(type(deliv).__name__ + ("(direct)" if getattr(deliv, "direct", False) else ""))
        pipe = PostHeadPipeline(dev_heads, STRIDES, kw, outs=deliv.outs, delivery=deliv)
        for rep, n_steps in enumerate((5, 4, 1, 2, 3, 8)):  # several runs: the flags / buffers must be reusable; 1 and 2 = the short-run graphs
            pipe.k, pipe.pending = 0, None
            if hasattr(deliv, "ctrl"):  # wipe rank 0's slots: a batch that is not delivered in THIS run must not pass
                deliv.local[: deliv.ctrl].zero_()
                torch.cuda.synchronize()
                dist.barrier()
            for k in range(n_steps):
                deliv.before_write((k - 1) & 1)
                done = pipe.step()
                if done is not None:
                    deliv.after_write(done)
            deliv.before_write((n_steps - 1) & 1)
            last = pipe.flush()
            deliv.after_write(last)
            deliv.drain()
            dist.barrier()
            torch.cuda.synchronize()
            got = deliv.result(last)
            streamed[(kinds[-1], rep)] = got
        del pipe, deliv
    if rank == 0:
        full_d, full_c = run(range(n_images))
        for key, (sd, sc) in streamed.items():
            assert torch.equal(sc, full_c) and torch.equal(sd, full_d), f"streamed delivery {key}: rank 0's buffer != single-GPU result"
        print("DELIVERIES", kinds)
        assert torch.equal(c, full_c) and torch.equal(d, full_d), "equal shards: gathered != single-GPU result"
        full_d2, full_c2 = run(range(n2))
        assert torch.equal(c2, full_c2) and torch.equal(d2, full_d2), "ragged shards: gathered != single-GPU result"
        print(f"MULTI_GPU_OK world={world} images={n_images}+{n2}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

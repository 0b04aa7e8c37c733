"""Sweep of the ark -> ark product path (bench.py's `product` block) over batch size and reader threads.
    python tools/bench_product.py [batch_frames,...] [threads,...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from xvector_b200 import synthetic  # noqa: E402


class A:
    topology = "ModelWithoutDropoutTdnn"


def main():
    frames = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "400000,200000,100000,50000").split(",")]
    threads = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4,8,16").split(",")]
    topo = bench.TOPOLOGIES[A.topology]
    params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], weight_set="B")
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=dev)
    for bf in frames:
        for th in threads:
            os.environ["XVEC_BATCH_FRAMES"] = str(bf)
            os.environ["XVEC_READER_THREADS"] = str(th)
            r = bench.measure_product(A, dev, rank, world, params, topo)
            if rank != 0:
                continue
            if "stream_output_variant" in r:
                v = r["stream_output_variant"]
                print(json.dumps(dict(stream_variant=True, value=v["value"], seconds=v["seconds"], breakdown=v["rank0_breakdown_s"])), flush=True)
            print(json.dumps(dict(batch_frames=bf, threads=th, value=r.get("value"), runs=r.get("all_runs_s"),
                                  breakdown=r.get("rank0_breakdown_s"), error=r.get("error"), trace=r.get("trace"))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

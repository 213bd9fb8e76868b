"""Stage IV: complementarity inference (README.md:14-16,30-32 of the reference names this entry point; the
script itself is absent there, the shipped logic is RankingLoss, codae/tool/metering.py:46-79).

For each query outfit: zero the requested slot, run the trained DAE, and rank the catalog items of that
category against the reconstructed slot (squared error by default, or cosine similarity), returning the
top-k.  The catalog is sharded by rows across ranks (torchrun), merged with an all-gather of k entries.

--mode swap scores every candidate SWAP by the reconstruction error of the whole swapped outfit instead
(candidate substituted into the slot, DAE forward per candidate; codae.tool.SwapScorer).
"""
import argparse
import json

import torch
import yaml

import _common  # noqa: F401
from _common import init_distributed, synthetic_categories
from codae.dataset import ConcatenatedEmbeddingDataset
from codae.model import EmbeddingDenoisingAutoencoder
from codae.tool import load_dataset_of_embeddings
from codae.tool.inference import ComplementarityScorer, SwapScorer, predict_slot, shard_rows


def parse():
    parser = argparse.ArgumentParser(description='Rank catalog items for a missing outfit slot.')
    parser.add_argument('--embedding_path', type=str, default=None)
    parser.add_argument('--config', type=str, required=True)
    parser.add_argument('--model_path', type=str, default=None)
    parser.add_argument('--slot', type=int, default=0)
    parser.add_argument('--k', type=int, default=10)
    parser.add_argument('--metric', type=str, default="sqerr", choices=["sqerr", "cosine"])
    parser.add_argument('--mode', type=str, default="slot", choices=["slot", "swap"])
    parser.add_argument('--compute_dtype', '--dtype', dest="compute_dtype", type=str, default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    parser.add_argument('--queries', type=int, default=4)
    parser.add_argument('--synthetic', type=int, default=0)
    parser.add_argument('--catalog_dtype', type=str, default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    return parser.parse_args()


if __name__ == "__main__":
    args = parse()
    rank, world, device = init_distributed()
    with open(args.config, 'r') as stream:
        config = yaml.safe_load(stream)
    cats = config["DATASET"]["USED_CATEGORY"]
    E = config["DATASET"]["EMBEDDING_SIZE"]
    if args.synthetic > 0:
        dataset = ConcatenatedEmbeddingDataset.from_tensors(synthetic_categories(args.synthetic, len(cats), E, config["SEED"]), cats)
    else:
        dataset = load_dataset_of_embeddings(embedding_path=args.embedding_path, config=config, cache_dir="tmp/")
    torch.manual_seed(config["SEED"])
    model = EmbeddingDenoisingAutoencoder(io_size=E * len(cats), z_size=config["MODEL"]["Z_SIZE"], embedding_size=E,
                                          nb_input_layer=config["MODEL"]["NB_INPUT_LAYER"],
                                          nb_output_layer=config["MODEL"]["NB_OUTPUT_LAYER"],
                                          steep_layer_size=config["MODEL"]["STEEP_LAYER_SIZE"])
    if args.model_path:
        model.load_state_dict(torch.load(args.model_path, map_location="cpu"))
    model.to(device).set_compute_dtype(args.compute_dtype)
    lo, n_local = shard_rows(dataset.nb_observation, world, rank)
    shard = dataset.data_per_category[args.slot][lo:lo + n_local].to(device)
    if args.catalog_dtype == "bf16":
        shard = shard.to(torch.bfloat16)
    outfits = dataset.data[:args.queries].to(device)
    if args.mode == "swap":
        if args.metric != "sqerr":
            raise Exception("Swap mode scores by squared reconstruction error.")
        scorer = SwapScorer(model, shard.contiguous(), E, k=args.k, inv_scale=1.0 / dataset.scale, row_offset=lo)
        res = [scorer.topk(outfits[q], args.slot) for q in range(outfits.shape[0])]
        if rank == 0:
            print(json.dumps({"slot": args.slot, "mode": "swap", "metric": "sqerr", "k": args.k,
                              "indices": [i.cpu().tolist() for _, i in res], "scores": [s.cpu().tolist() for s, _ in res]}))
        raise SystemExit(0)
    scorer = ComplementarityScorer(shard.contiguous(), E, metric=args.metric, k=args.k, inv_scale=1.0 / dataset.scale,
                                   row_offset=lo)
    p = predict_slot(model, outfits, args.slot, E)
    scores, idx = scorer.topk(p)
    if rank == 0:
        print(json.dumps({"slot": args.slot, "metric": args.metric, "k": args.k,
                          "indices": idx.cpu().tolist(), "scores": scores.cpu().tolist()}))

"""Train the co-occurrence denoising autoencoder on concatenated per-category embeddings.

Same flags and YAML keys as the reference's script/train_dae_on_embedding.py; the loop body
(reference :194-223 training, :241-261 validation) runs as FusedStep's kernel sequence on a B200.
Extra optional flags: --synthetic N (generate N synthetic observations instead of reading --embedding_path),
--graph (capture the step into a CUDA graph).  Launch with torchrun for data parallelism.
"""
import argparse
import logging
import math
import os

import numpy as np
import torch
import yaml

import _common  # noqa: F401  (sets sys.path)
from _common import epoch_batches, init_distributed, synthetic_categories
from codae.dataset import ConcatenatedEmbeddingDataset
from codae.model import EmbeddingDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep, RankingLoss, display_info, get_date, load_dataset_of_embeddings, set_logging


def parse():
    parser = argparse.ArgumentParser(description='Train denoising autoencoder.')
    parser.add_argument('--embedding_path', type=str, required=False, default=None)
    parser.add_argument('--output_path', type=str, required=True)
    parser.add_argument('--config', type=str, required=True)
    parser.add_argument('--debug', type=bool, default=False)
    parser.add_argument('--rank', type=bool, default=False)
    parser.add_argument('--nb_missing', type=int, default=1)
    parser.add_argument('--synthetic', type=int, default=0)
    parser.add_argument('--graph', action="store_true",
                        help="device-side sampler + CUDA graphs: one graph launch per 32 training steps (FusedStep.train_epoch)")
    parser.add_argument('--epochs', type=int, default=None)
    parser.add_argument('--dtype', type=str, default=None, choices=["fp32", "bf16", "fp32_simt"],
                        help="compute engine: fp32 (reference precision) or bf16 tensor cores; overrides MODEL.DTYPE of the config")
    return parser.parse_args()


if __name__ == "__main__":
    args = parse()
    rank, world, device = init_distributed()
    log = set_logging(logging_level=(logging.DEBUG if args.debug else logging.INFO), log_file_path="log/" if rank == 0 else None)
    with open(args.config, 'r') as stream:
        config = yaml.safe_load(stream)
    cats = config["DATASET"]["USED_CATEGORY"]
    E = config["DATASET"]["EMBEDDING_SIZE"]

    if args.synthetic > 0:
        dataset = ConcatenatedEmbeddingDataset.from_tensors(synthetic_categories(args.synthetic, len(cats), E, config["SEED"]), cats)
    else:
        dataset = load_dataset_of_embeddings(embedding_path=args.embedding_path, config=config, cache_dir="tmp/")
    log.info("Dataset STD = " + str(torch.std(dataset.data)))

    indices = list(range(dataset.nb_observation))
    nb_train = math.floor(dataset.nb_observation * config["DATASET"]["SPLIT"][0])
    nb_validation = dataset.nb_observation - nb_train
    if config["DATASET"]["SHUFFLE"]:
        np.random.seed(config["SEED"])
        np.random.shuffle(indices)
    train_indices, validation_indices = indices[:nb_train], indices[nb_train:]

    io_size = E * len(cats)
    torch.manual_seed(config["SEED"])   # identical initial weights on every rank
    model = EmbeddingDenoisingAutoencoder(io_size=io_size, z_size=config["MODEL"]["Z_SIZE"], embedding_size=E,
                                          nb_input_layer=config["MODEL"]["NB_INPUT_LAYER"],
                                          nb_output_layer=config["MODEL"]["NB_OUTPUT_LAYER"],
                                          steep_layer_size=config["MODEL"]["STEEP_LAYER_SIZE"])
    model.set_compute_dtype(args.dtype or config["MODEL"].get("DTYPE", "fp32"))
    model.to(device)
    dataset.to(device)
    corrupter = Corrupter(nb_observation=dataset.nb_observation, arch=dataset.arch, k_max=args.nb_missing, device=device,
                          seed=config["SEED"])
    if rank == 0:
        display_info(config, dataset.nb_observation, {})
        log.info(model)
    trainer = FusedStep(model, corrupter, dataset.data, lr=config["MODEL"]["LEARNING_RATE"],
                        weight_decay=config["MODEL"]["WEIGHT_DECAY"],
                        clip=bool(config["MODEL"].get("TRUNK_GRAD", False)),   # key missing in the modanet yaml
                        world_size=world, use_graph=args.graph)
    ranking_loss = RankingLoss(dataset, validation_indices, device=device)
    book = {"ftl": [], "ptl": [], "fvl": [], "pvl": [], "rl": []}
    B = config["MODEL"]["BATCH_SIZE"]
    rng = np.random.RandomState(config["SEED"])
    S = dataset.nb_used_category
    train_idx_dev = torch.as_tensor(train_indices, dtype=torch.int64, device=device)
    sampler_gen = torch.Generator(device=device)
    sampler_gen.manual_seed(config["SEED"])

    for epoch in range(args.epochs or config["MODEL"]["EPOCH"]):
        if rank == 0:
            log.info("===================================================== EPOCH = %d" % epoch)
        trainer.reset_monitors()
        if args.graph:
            # sampler on the device, 32 steps per CUDA-graph launch (FusedStep.train_epoch); every rank draws the same permutation
            trainer.train_epoch(train_idx_dev, B, generator=sampler_gen, rank=rank)
        else:
            for local_idx, global_b in epoch_batches(train_indices, B, rng, rank, world):
                idx = torch.as_tensor(local_idx, dtype=torch.int64).pin_memory()
                trainer.step(idx, run=0, global_batch=global_b)
        mon = trainer.read_monitors()                      # one D2H per epoch instead of two per step
        acc = torch.tensor([mon["full"], mon["partial"]], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(acc)
        ftl = math.sqrt(acc[0].item() / (dataset.nb_predictor * nb_train))
        ptl = math.sqrt(acc[1].item() / (nb_train * dataset.nb_predictor / S))
        book["ftl"].append(ftl)
        book["ptl"].append(ptl)

        trainer.reset_monitors()
        rl = torch.zeros((), dtype=torch.float64, device=device)       # accumulated on the device: one read-back per epoch
        for local_idx, _ in epoch_batches(validation_indices, B, rng, rank, world):
            if len(local_idx) == 0:
                continue
            idx = torch.as_tensor(local_idx, dtype=torch.int64, device=device)
            out = trainer.evaluate(idx, run=0)
            if args.rank and args.nb_missing == 1:
                _, fmask = corrupter.get_masks(idx, 0)
                rl += ranking_loss.get_tensor(out, fmask, idx)
        mon = trainer.read_monitors()
        acc = torch.cat([torch.tensor([mon["full"], mon["partial"]], dtype=torch.float64, device=device), rl.view(1)])
        if world > 1:
            torch.distributed.all_reduce(acc)
        fvl = math.sqrt(acc[0].item() / (dataset.nb_predictor * nb_validation))
        pvl = math.sqrt(acc[1].item() / (nb_validation * dataset.nb_predictor / S))
        book["fvl"].append(fvl)
        book["pvl"].append(pvl)
        book["rl"].append(acc[2].item() / nb_validation)
        if rank == 0:
            log.info("TRAINING FULL ERROR      = %7f" % ftl)
            log.info("TRAINING PARTIAL ERROR   = %7f" % ptl)
            log.info("VALIDATION FULL ERROR    = %7f" % fvl)
            log.info("VALIDATION PARTIAL ERROR = %7f" % pvl)
            log.info("VALIDATION RANKING ERROR = %7f" % book["rl"][-1])

    trainer.flush()        # data parallel: collective gather of the sharded fp32 master weights before they are read; no-op on one GPU
    if rank == 0:
        log.info("TRAINING HAS ENDED.")
        d = os.path.join(args.output_path, get_date() + "_train_" + config["DATASET"]["NAME"])
        os.makedirs(d, exist_ok=True)
        np.savez(os.path.join(d, "metrics.npz"), **{k: np.asarray(v) for k, v in book.items()})
        torch.save({k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
                   os.path.join(d, "model.pt"))   # new: stage IV needs a --model_path (clones: params are views of one flat buffer)
        log.info("Data saved in directory %s" % d)

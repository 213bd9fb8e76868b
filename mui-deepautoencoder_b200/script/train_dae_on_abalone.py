"""Train the mixed-variable denoising autoencoder on the UCI abalone table.

Same flags and YAML keys as the reference's script/train_dae_on_abalone.py; the loop body (reference :200-236
training with nb_run augmentation passes, :276-301 validation) runs as FusedStep's kernel sequence.
Extra optional flag: --synthetic N (a synthetic table of the abalone shape: C3 + 8 x R1).
"""
import argparse
import logging
import math
import os

import numpy as np
import torch
import yaml

import _common  # noqa: F401
from _common import epoch_batches, init_distributed
from codae.dataset import MixedVariableDataset
from codae.model import MixedVariableDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep, Normalizer, get_date, set_logging

NAMES = ["Sex", "Length", "Diameter", "Height", "Whole", "Shucked", "Viscera", "Shell", "Rings"]


def parse():
    parser = argparse.ArgumentParser(description='Train denoising autoencoder.')
    parser.add_argument('--dataset_path', type=str, required=False, default=None)
    parser.add_argument('--output_path', type=str, default="output/")
    parser.add_argument('--config', type=str, required=True)
    parser.add_argument('--debug', type=bool, default=False)
    parser.add_argument('--nb_missing', type=int, default=1)
    parser.add_argument('--synthetic', type=int, default=0)
    parser.add_argument('--epochs', type=int, default=None)
    parser.add_argument('--dtype', type=str, default="fp32", choices=["fp32"],
                        help="tabular widths (11 x 11 layers) run on the exact-fp32 engine only; the flag exists for symmetry")
    return parser.parse_args()


class _MinMax:
    def __init__(self, a):
        self.data_min_, self.data_max_ = a.min(axis=0), a.max(axis=0)
        self.data_range_ = self.data_max_ - self.data_min_

    def transform(self, a):
        return (a - self.data_min_) / self.data_range_


def load_frame(args, seed):
    import pandas as pd
    if args.synthetic > 0:
        r = np.random.RandomState(seed)
        cols = {"Sex": r.choice(["M", "F", "I"], size=args.synthetic)}
        for n in NAMES[1:]:
            cols[n] = r.rand(args.synthetic) * r.uniform(0.5, 30.0)
        return pd.DataFrame(cols)
    with open(os.path.join(args.dataset_path, "abalone.data"), 'r') as f:
        return pd.read_csv(f, sep=",", header=0, names=NAMES)   # header=0 like the reference (:85)


if __name__ == "__main__":
    args = parse()
    rank, world, device = init_distributed()
    log = set_logging(logging_level=(logging.DEBUG if args.debug else logging.INFO), log_file_path="log/" if rank == 0 else None)
    with open(args.config, 'r') as stream:
        config = yaml.safe_load(stream)
    frame = load_frame(args, config["SEED"])
    num = frame.iloc[:, 1:].to_numpy(dtype=np.float64)      # cast first: pandas >= 2 refuses the in-place int->float write
    scaler = _MinMax(num)
    frame = frame.astype({c: np.float64 for c in frame.columns[1:]})
    frame.iloc[:, 1:] = scaler.transform(num)
    tensor_normazer = Normalizer(normalizer=scaler, device=device)
    dataset = MixedVariableDataset(frame)

    indices = list(range(dataset.nb_observation))
    nb_train = math.floor(dataset.nb_observation * config["DATASET"]["SPLIT"][0])
    nb_validation = dataset.nb_observation - nb_train
    if config["DATASET"]["SHUFFLE"]:
        np.random.seed(config["SEED"])
        np.random.shuffle(indices)
    train_indices, validation_indices = indices[:nb_train], indices[nb_train:]

    corrupter = Corrupter(nb_observation=dataset.nb_observation, arch=dataset.arch, k_max=args.nb_missing, device=device,
                          seed=config["SEED"])
    torch.manual_seed(config["SEED"])
    model = MixedVariableDenoisingAutoencoder(arch=dataset.arch, io_size=dataset.io_size, z_size=config["MODEL"]["Z_SIZE"],
                                              device=device, nb_input_layer=config["MODEL"]["NB_INPUT_LAYER"],
                                              nb_output_layer=config["MODEL"]["NB_OUTPUT_LAYER"],
                                              steep_layer_size=config["MODEL"]["STEEP_LAYER_SIZE"])
    model.to(device)
    dataset.to(device)
    first_num = dataset.arch[1]["position"]
    trainer = FusedStep(model, corrupter, dataset.data, lr=config["MODEL"]["LEARNING_RATE"],
                        weight_decay=config["MODEL"]["WEIGHT_DECAY"], clip=bool(config["MODEL"].get("TRUNK_GRAD", False)),
                        world_size=1,   # 792 parameters: replicas only, data parallelism is pointless here
                        mixed=dict(arch=dataset.arch, weight=[0.4] + [1] * (len(dataset.arch) - 1),
                                   norm_scale=tensor_normazer.scale, norm_min=tensor_normazer.min, norm_first=first_num))
    B = config["MODEL"]["BATCH_SIZE"]
    rng = np.random.RandomState(config["SEED"])
    K, V = args.nb_missing, len(dataset.arch)
    per_k = corrupter.nb_corruption_per_k

    def finish(mon, n):
        f_k, p_k = mon["ftl_per_k"].copy(), mon["ptl_per_k"].copy()
        for i in range(K):
            f_k[i, :] /= n * sum(per_k[:i + 1])
            p_k[i, :] /= n * sum(per_k[:i + 1]) / dataset.nb_predictor
        f = math.sqrt(mon["ftl"] / (sum(per_k) * n))
        p = math.sqrt(mon["ptl"] / (sum(per_k) * n / dataset.nb_predictor))
        f_k[:, 1:], p_k[:, 1:] = np.sqrt(f_k[:, 1:]), np.sqrt(p_k[:, 1:])
        return f, p, f_k, p_k

    for epoch in range(args.epochs or config["MODEL"]["EPOCH"]):
        log.info("===================================================== EPOCH = %d" % epoch)
        trainer.reset_monitors()
        for run in range(corrupter.nb_run):
            for local_idx, gb in epoch_batches(train_indices, B, rng):
                trainer.step(torch.as_tensor(local_idx, dtype=torch.int64, device=device), run=run, global_batch=gb)
        ftl, ptl, ftl_k, ptl_k = finish(trainer.read_monitors(), nb_train)
        log.info("TRAINING PARTIAL ERROR = %7f" % np.mean(ptl_k))
        trainer.reset_monitors()
        for run in range(corrupter.nb_run):
            for local_idx, gb in epoch_batches(validation_indices, B, rng):
                trainer.evaluate(torch.as_tensor(local_idx, dtype=torch.int64, device=device), run=run)
        fvl, pvl, fvl_k, pvl_k = finish(trainer.read_monitors(), nb_validation)
        log.info("VALIDATION PARTIAL ERROR = %7f" % np.mean(pvl_k))
        for i in range(K):
            log.info("k=%d  " % (i + 1) + " ".join("%s=%f" % (n, pvl_k[i][j]) for j, n in enumerate(dataset.variable_names)))
    log.info("TRAINING HAS ENDED.")

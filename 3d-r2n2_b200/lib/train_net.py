"""Training entry of the reference (lib/train_net.py:15-85): build the network and solver, start the data
workers on two queues, train, and always stop the workers.  Changes in the wiring: the trainer reads the
queues through DeviceBatchQueue, which turns the workers' raw batches into device tensors ahead of the step
(lib/data_process.py); and when the process was started by a multi-process launcher (torchrun, one process per
GPU) the same function trains data-parallel: every rank takes a disjoint share of `category_model_id_pair`,
all ranks start from rank 0's parameters, the number of views per step is one shared schedule
(dist.SharedViewCountQueue; reference draw: lib/data_process.py:127-130), gradients are all-reduced over NCCL
(dist.GradSync) and only rank 0 prints and saves."""
from multiprocessing import Queue

import numpy as np

from ..models import load_model
from .config import cfg
from .data_io import category_model_id_pair
from .data_process import DeviceBatchQueue, kill_processes, make_data_processes
from .solver import Solver

# module-level like the reference, so that an interrupted run can still be cleaned up
train_queue, val_queue, train_processes, val_processes = None, None, None, None


def _stop_workers():
    for q, procs in ((train_queue, train_processes), (val_queue, val_processes)):
        if q is not None and procs:
            kill_processes(q, procs)


def train_net():
    '''Main training function'''
    global train_queue, val_queue, train_processes, val_processes
    NetClass = load_model(cfg.CONST.NETWORK_CLASS)
    if NetClass is None:
        raise ValueError('unknown network class %r' % cfg.CONST.NETWORK_CLASS)
    from .. import dist as r2dist
    rank, world, _ = r2dist.init_from_env()
    net = NetClass()
    if net.is_x_tensor4 and cfg.CONST.N_VIEWS > 1:
        raise ValueError('Do not set the config.CONST.N_VIEWS > 1 when using'
                         'single-view reconstruction network')
    solver = Solver(net)
    if world > 1:
        r2dist.broadcast_params(net.engine)
        r2dist.attach(solver)
        # per-rank worker randomness (mini-batch order, rendering ids, augmentation): main.py:85 seeds numpy once
        np.random.seed((cfg.CONST.RNG_SEED + 1000003 * rank) % (2 ** 32))

    train_queue = Queue(cfg.QUEUE_SIZE)
    val_queue = Queue(cfg.QUEUE_SIZE)
    dev_train = dev_val = None
    random_views = cfg.TRAIN.RANDOM_NUM_VIEWS
    try:
        train_pairs = r2dist.shard(category_model_id_pair(dataset_portion=cfg.TRAIN.DATASET_PORTION), rank, world)
        val_pairs = r2dist.shard(category_model_id_pair(dataset_portion=cfg.TEST.DATASET_PORTION), rank, world)
        if world > 1:
            cfg.TRAIN.RANDOM_NUM_VIEWS = False      # read by the forked workers: always N_VIEWS renders per sample
        try:
            train_processes = make_data_processes(train_queue, train_pairs, cfg.TRAIN.NUM_WORKER, repeat=True)
            val_processes = make_data_processes(val_queue, val_pairs, 1, repeat=True, train=False)
        finally:
            cfg.TRAIN.RANDOM_NUM_VIEWS = random_views
        src_train = train_queue
        if world > 1:
            seed = r2dist.shared_seed(cfg.CONST.RNG_SEED)
            src_train = r2dist.SharedViewCountQueue(train_queue, seed, cfg.CONST.N_VIEWS, random_views)
        dev_train = DeviceBatchQueue(src_train)
        dev_val = DeviceBatchQueue(val_queue, depth=1)
        return solver.train(dev_train, dev_val)
    finally:
        for d in (dev_train, dev_val):
            if d is not None:
                d.close()
        print('Wait until the dataprocesses to end')
        _stop_workers()
        train_processes, val_processes = None, None

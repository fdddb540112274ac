"""Training entry of the reference (lib/train_net.py:15-85): build the network and solver, start the data
workers on two queues, train, and always stop the workers.  The only change in the wiring: the trainer reads
the queues through DeviceBatchQueue, which turns the workers' raw batches into device tensors ahead of the
step (lib/data_process.py)."""
from multiprocessing import Queue

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
    net = NetClass()
    if net.is_x_tensor4 and cfg.CONST.N_VIEWS > 1:
        raise ValueError('Do not set the config.CONST.N_VIEWS > 1 when using'
                         'single-view reconstruction network')
    solver = Solver(net)

    train_queue = Queue(cfg.QUEUE_SIZE)
    val_queue = Queue(cfg.QUEUE_SIZE)
    dev_train = dev_val = None
    try:
        train_processes = make_data_processes(
            train_queue, category_model_id_pair(dataset_portion=cfg.TRAIN.DATASET_PORTION),
            cfg.TRAIN.NUM_WORKER, repeat=True)
        val_processes = make_data_processes(
            val_queue, category_model_id_pair(dataset_portion=cfg.TEST.DATASET_PORTION), 1,
            repeat=True, train=False)
        dev_train = DeviceBatchQueue(train_queue)
        dev_val = DeviceBatchQueue(val_queue, depth=1)
        return solver.train(dev_train, dev_val)
    finally:
        for d in (dev_train, dev_val):
            if d is not None:
                d.close()
        print('Wait until the dataprocesses to end')
        _stop_workers()
        train_processes, val_processes = None, None

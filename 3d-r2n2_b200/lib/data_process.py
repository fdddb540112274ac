"""Input pipeline: the worker processes and queue interface of lib/data_process.py:32-209,
re-cut for a step that takes ~12 ms.

The reference's workers decode every render with PIL, composite / crop / flip / scale it in numpy and push
float32 batches (44 MB at 36 x 5 views) through an mp.Queue; the trainer then copies them to the device.
Here the workers only decode: they put the raw uint8 RGBA renders (13.5 MB per batch), the random
augmentation parameters -- drawn from numpy's global RNG in the reference's order, so a seeded worker makes
the same choices -- and the byte occupancy grid on the queue.  `DeviceBatchQueue` wraps that queue on the
trainer's side: a thread stages each batch in pinned memory, copies it on a side stream and runs
r2n2_preprocess_views / r2n2_voxel_labels there, so `get()` hands `Solver.train` device tensors
(x (T,B,3,H,W), y (B,n_vox,2,n_vox,n_vox)) that are already resident when the step starts.

Class and function names, constructor arguments and the shutdown protocol are the reference's."""
import ctypes
import os
import queue as _queue
import sys
import threading
import time
import traceback
from multiprocessing import Event, Process

import numpy as np

from .binvox_rw import read_as_3d_array
from .config import cfg
from .data_augmentation import draw_params
from .data_io import get_rendering_file, get_voxel_file


def print_error(func):
    '''Flush out error messages. Mainly used for debugging separate processes (lib/data_process.py:19-29).

    In a forked worker the function never returns: the trainer has usually initialised CUDA before it starts
    the workers, so the child carries a copy of the parent's CUDA state that must not be torn down by the
    interpreter's exit handlers (they ran driver calls on inherited handles while the parent was capturing a
    CUDA graph and invalidated the capture).  The child flushes its queue and leaves through os._exit.'''

    def func_wrapper(self, *args, **kwargs):
        try:
            return func(self, *args, **kwargs)
        except Exception:
            traceback.print_exception(*sys.exc_info())
            sys.stdout.flush()
        finally:
            if getattr(self, '_parent_pid', None) not in (None, os.getpid()):      # running as the child process
                try:
                    sys.stdout.flush()
                    q = self.data_queue
                    if hasattr(q, 'close') and hasattr(q, 'join_thread'):
                        q.close()
                        q.join_thread()                                            # everything put() reaches the pipe
                finally:
                    os._exit(0)

    return func_wrapper


class RawViews(object):
    """The image half of a raw batch: rgba uint8 (T, B, Hs, Ws, 4) and params int32 (T, B, 6) =
    {crop row, crop column, flip, bg r, bg g, bg b} per render (the arguments of r2n2_preprocess_views).
    Slots of a short last batch keep alpha 0 and a black background: they come out as zeros, like the
    untouched rows of the reference's np.zeros batch."""
    __slots__ = ('rgba', 'params')

    def __init__(self, rgba, params):
        self.rgba = rgba
        self.params = params

    def __getstate__(self):
        return (self.rgba, self.params)

    def __setstate__(self, s):
        self.rgba, self.params = s

    @property
    def shape(self):
        """Shape of the float batch this becomes: (T, B, 3, IMG_H, IMG_W)."""
        return (self.rgba.shape[0], self.rgba.shape[1], 3, cfg.CONST.IMG_H, cfg.CONST.IMG_W)


class DataProcess(Process):
    """lib/data_process.py:32-97."""

    def __init__(self, data_queue, data_paths, repeat=True):
        '''
        data_queue : Multiprocessing queue
        data_paths : list of data and label pair used to load data
        repeat : if set True, return data until exit is set
        '''
        super(DataProcess, self).__init__()
        self.data_queue = data_queue
        self.data_paths = data_paths
        self.num_data = len(data_paths)
        self.repeat = repeat
        self.batch_size = cfg.CONST.BATCH_SIZE
        self.exit = Event()
        self.shuffle_db_inds()

    def shuffle_db_inds(self):
        if self.repeat:
            self.perm = np.random.permutation(np.arange(self.num_data))
        else:
            self.perm = np.arange(self.num_data)
        self.cur = 0

    def get_next_minibatch(self):
        if (self.cur + self.batch_size) >= self.num_data and self.repeat:
            self.shuffle_db_inds()
        db_inds = self.perm[self.cur:min(self.cur + self.batch_size, self.num_data)]
        self.cur += self.batch_size
        return db_inds

    def shutdown(self):
        self.exit.set()

    @print_error
    def run(self):
        while not self.exit.is_set() and self.cur <= self.num_data:
            db_inds = self.get_next_minibatch()
            data_list = [self.load_datum(self.data_paths[i]) for i in db_inds]
            label_list = [self.load_label(self.data_paths[i]) for i in db_inds]
            batch_data = np.array(data_list).astype(np.float32)
            batch_label = np.array(label_list).astype(np.float32)
            self.data_queue.put((batch_data, batch_label), block=True)

    def load_datum(self, path):
        pass

    def load_label(self, path):
        pass


class ReconstructionDataProcess(DataProcess):
    """lib/data_process.py:100-173, emitting raw batches: (RawViews, occupancy uint8 (B, n_vox, n_vox, n_vox):
    0 empty, 1 occupied, 2 in the unfilled slots of a short last batch)."""

    def __init__(self, data_queue, category_model_pair, background_imgs=[], repeat=True, train=True):
        self.repeat = repeat
        self.train = train
        self.background_imgs = background_imgs
        super(ReconstructionDataProcess, self).__init__(data_queue, category_model_pair, repeat=repeat)

    def make_batch(self):
        """One pass of the reference's loop body (lib/data_process.py:121-156), random draws in its order:
        the mini-batch indices, the number of views, then per sample the rendering ids and per render the
        background colour, crop and flip (preprocess_img), then the label."""
        n_vox = cfg.CONST.N_VOX
        n_views = cfg.CONST.N_VIEWS
        db_inds = self.get_next_minibatch()
        if cfg.TRAIN.RANDOM_NUM_VIEWS:
            curr_n_views = np.random.randint(n_views) + 1
        else:
            curr_n_views = n_views
        rgba = params = None
        occ = np.full((self.batch_size, n_vox, n_vox, n_vox), 2, np.uint8)     # 2 = slot not filled
        for batch_id, db_ind in enumerate(db_inds):
            category, model_id = self.data_paths[db_ind]
            image_ids = np.random.choice(cfg.TRAIN.NUM_RENDERING, curr_n_views)
            for view_id, image_id in enumerate(image_ids):
                im, p = self.load_img(category, model_id, image_id)
                if rgba is None:
                    rgba = np.zeros((curr_n_views, self.batch_size) + im.shape, np.uint8)
                    params = np.zeros((curr_n_views, self.batch_size, 6), np.int32)
                elif im.shape != rgba.shape[2:]:
                    raise ValueError('render %s/%s/%02d is %s, the batch holds %s images' %
                                     (category, model_id, image_id, im.shape[:2], rgba.shape[2:4]))
                rgba[view_id, batch_id] = im
                params[view_id, batch_id] = p
            voxel = self.load_label(category, model_id)
            if voxel.data.shape != (n_vox, n_vox, n_vox):
                raise ValueError('label %s/%s is %s, expected %d^3' % (category, model_id, voxel.data.shape, n_vox))
            occ[batch_id] = voxel.data
        if rgba is None:                                          # empty tail batch
            rgba = np.zeros((curr_n_views, self.batch_size, cfg.CONST.IMG_H, cfg.CONST.IMG_W, 4), np.uint8)
            params = np.zeros((curr_n_views, self.batch_size, 6), np.int32)
        return RawViews(rgba, params), occ

    @print_error
    def run(self):
        while not self.exit.is_set() and self.cur <= self.num_data:
            self.data_queue.put(self.make_batch(), block=True)
        print('Exiting')

    def load_img(self, category, model_id, image_id):
        """The decoded render as uint8 (Hs, Ws, 4) (an RGB file gets alpha 255: nothing is composited, as in
        add_random_color_background, lib/data_augmentation.py:41-54) and its six augmentation parameters."""
        from PIL import Image
        im = np.array(Image.open(get_rendering_file(category, model_id, image_id)))
        if im.ndim != 3 or im.shape[2] not in (3, 4) or im.dtype != np.uint8:
            raise ValueError('expected an 8-bit RGB(A) render, got %s %s' % (im.dtype, im.shape))
        if im.shape[2] == 3:
            im = np.concatenate([im, np.full(im.shape[:2] + (1,), 255, np.uint8)], axis=2)
        return im, draw_params(1, self.train, im.shape[:2])[0]

    def load_label(self, category, model_id):
        with open(get_voxel_file(category, model_id), 'rb') as f:
            return read_as_3d_array(f)


def kill_processes(queue, processes):
    """lib/data_process.py:176-189."""
    print('Signal processes')
    for p in processes:
        p.shutdown()
    print('Empty queue')
    while not queue.empty():
        time.sleep(0.5)
        queue.get(False)
    print('kill processes')
    for p in processes:
        p.terminate()


def make_data_processes(queue, data_paths, num_workers, repeat=True, train=True):
    '''Make a set of data processes for parallel data loading (lib/data_process.py:192-202).'''
    processes = []
    for _ in range(num_workers):
        process = ReconstructionDataProcess(queue, data_paths, repeat=repeat, train=train)
        process.start()
        processes.append(process)
    return processes


def get_while_running(data_process, data_queue, sleep_time=0):
    """lib/data_process.py:205-216: drain the queue until the worker has exited."""
    while True:
        time.sleep(sleep_time)
        try:
            batch_data, batch_label = data_queue.get_nowait()
        except _queue.Empty:
            if not data_process.is_alive():
                break
            else:
                continue
        yield batch_data, batch_label


# --------------------------------------------------------------------------------------- trainer side
def batch_to_device(batch_img, batch_voxel, stream=None, staging=None, out_dtype=None):
    """One queue item -> (x, y) CUDA tensors.  A raw batch goes through the two kernels; a float batch (what a
    reference-style worker would send) is only copied.  `staging` = (pinned uint8 buffer for the renders,
    pinned int32 buffer for the parameters, pinned uint8 buffer for the grid) or None."""
    import torch
    from .. import ops
    out_dtype = out_dtype or torch.float32
    if not torch.cuda.is_available():
        raise RuntimeError('the input pipeline runs its kernels on the GPU: no CUDA device')
    stream = stream or torch.cuda.current_stream()
    with torch.cuda.stream(stream):
        if isinstance(batch_img, RawViews):
            T, B, Hs, Ws, _ = batch_img.rgba.shape
            n = T * B
            if staging is not None:
                pr, pp, po = staging
                pr = pr[:batch_img.rgba.size].view(n, Hs, Ws, 4)
                pp = pp[:n * 6].view(n, 6)
                po = po[:batch_voxel.size].view(batch_voxel.shape)
                # one plain memcpy each, outside the GIL (ctypes releases it): torch's copy_ fans the 13.5 MB out
                # over every host core in every rank's staging thread, and a numpy copy holds the GIL for
                # milliseconds, which convoys the launching thread (measured: 2970 -> 1230 samples/s)
                for dst_t, src in ((pr, batch_img.rgba), (pp, batch_img.params),
                                   (po, np.asarray(batch_voxel, dtype=np.uint8))):
                    src = np.ascontiguousarray(src)
                    ctypes.memmove(dst_t.data_ptr(), src.ctypes.data, src.nbytes)
            else:
                pr = torch.from_numpy(batch_img.rgba).view(n, Hs, Ws, 4)
                pp = torch.from_numpy(batch_img.params).view(n, 6)
                po = torch.from_numpy(np.ascontiguousarray(batch_voxel, dtype=np.uint8))
            H, W = cfg.CONST.IMG_H, cfg.CONST.IMG_W
            p = batch_img.params.reshape(n, 6)
            if (p[:, :2].min() < 0 or int(p[:, 0].max()) + H > Hs or int(p[:, 1].max()) + W > Ws):
                raise ValueError('crop window outside the source image')
            x = torch.empty(n, 3, H, W, dtype=out_dtype, device='cuda')
            ops.preprocess_views(pr.cuda(non_blocking=True), pp.cuda(non_blocking=True), x)
            y = ops.voxel_labels(po.cuda(non_blocking=True))
            return x.view(T, B, 3, H, W), y
        x = torch.as_tensor(np.asarray(batch_img, dtype=np.float32)).cuda(non_blocking=True)
        y = torch.as_tensor(np.asarray(batch_voxel, dtype=np.float32)).cuda(non_blocking=True)
        return x, y


class DeviceBatchQueue(object):
    """`get()`-compatible wrapper of the workers' queue for `Solver.train(train_queue)` (lib/solver.py:111-186
    only calls `train_queue.get()`).  A daemon thread keeps up to `depth` batches converted and resident: the
    H2D copy of batch i+1 and its two kernels run on a side stream while step i computes."""

    def __init__(self, data_queue, depth=2, out_dtype=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('DeviceBatchQueue needs a CUDA device')
        self.data_queue = data_queue
        self.depth = max(1, int(depth))
        self.out_dtype = out_dtype
        self._ready = _queue.Queue(self.depth)
        self._stop = threading.Event()
        self._device = torch.cuda.current_device()
        self._stream = torch.cuda.Stream()
        self._slots = []                      # (pinned buffers, event of the last copy out of them)
        self._error = None
        self._thread = threading.Thread(target=self._work, name='r2n2-device-batch-queue', daemon=True)
        self._thread.start()

    def _staging(self, item):
        import torch
        batch_img, batch_voxel = item
        if not isinstance(batch_img, RawViews):
            return None, None
        T, B, Hs, Ws, _ = batch_img.rgba.shape
        need = (max(T, cfg.CONST.N_VIEWS) * B * Hs * Ws * 4, max(T, cfg.CONST.N_VIEWS) * B * 6, batch_voxel.size)
        def alloc():
            return (torch.empty(need[0], dtype=torch.uint8).pin_memory(),
                    torch.empty(need[1], dtype=torch.int32).pin_memory(),
                    torch.empty(need[2], dtype=torch.uint8).pin_memory())
        if not self._slots:
            # the whole ring at once, on the first batch: page-locking 13.5 MB takes milliseconds and must not
            # land in the middle of later steps
            self._slots = [[alloc(), None] for _ in range(self.depth + 1)]
        slot = self._slots.pop(0)
        self._slots.append(slot)
        if slot[1] is not None:
            slot[1].synchronize()             # the copies out of this pinned slot have finished
        if any(b.numel() < n for b, n in zip(slot[0], need)):
            slot[0] = alloc()
        return slot[0], slot

    def _work(self):
        import torch
        try:
            torch.cuda.set_device(self._device)
            while not self._stop.is_set():
                try:
                    item = self.data_queue.get(timeout=0.2)
                except _queue.Empty:
                    continue
                bufs, slot = self._staging(item)
                x, y = batch_to_device(item[0], item[1], self._stream, bufs, self.out_dtype)
                ev = torch.cuda.Event()
                ev.record(self._stream)
                if slot is not None:
                    slot[1] = ev
                while not self._stop.is_set():
                    try:
                        self._ready.put((x, y, ev), timeout=0.2)
                        break
                    except _queue.Full:
                        continue
        except Exception as e:                                     # surfaced by get()
            self._error = e
            traceback.print_exc()

    def get(self, block=True, timeout=None):
        import torch
        deadline = None if timeout is None else time.time() + timeout
        while True:
            if self._error is not None:
                raise RuntimeError('input pipeline thread failed') from self._error
            try:
                x, y, ev = self._ready.get(block=block, timeout=0.2 if block else None)
                break
            except _queue.Empty:
                if not block or (deadline is not None and time.time() > deadline):
                    raise
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        x.record_stream(cur)
        y.record_stream(cur)
        return x, y

    def empty(self):
        return self._ready.empty() and self.data_queue.empty()

    def close(self):
        self._stop.set()
        self._thread.join(timeout=5)

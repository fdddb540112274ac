"""Minimal mirror of lib/config.py: the global `cfg` with the reference's defaults for the keys
the hot path reads (lib/config.py:14-101) plus `cfg_from_list` (lib/config.py:137-161).
easydict is not a dependency; `_AttrDict` gives the same attribute access."""
from ast import literal_eval


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


__C = _AttrDict()
cfg = __C
__C.NET_NAME = 'res_gru_net'
__C.PROFILE = False

__C.CONST = _AttrDict()
__C.CONST.DEVICE = 'gpu0'
__C.CONST.RNG_SEED = 0
__C.CONST.IMG_W = 127
__C.CONST.IMG_H = 127
__C.CONST.N_VOX = 32
__C.CONST.N_VIEWS = 5
__C.CONST.BATCH_SIZE = 36
__C.CONST.NETWORK_CLASS = 'ResidualGRUNet'
__C.CONST.WEIGHTS = ''
# not in the reference: arithmetic of the dense contractions.
#   'fp32' -> exact fp32 FFMA kernels (parity path); 'bf16' -> tcgen05 bf16 operands, fp32 accumulate
__C.CONST.COMPUTE = 'fp32'
__C.CONST.OVERLAP_H2D = True   # new key: training inputs are copied view by view on a side stream, behind the first layers
__C.CONST.INFER_GRAPH = True   # new key: replay inference (Solver.test_output) as one CUDA graph per (views, has-y)

__C.DATASET = './experiments/dataset/shapenet_1000.json'       # lib/config.py:12
__C.QUEUE_SIZE = 15                                            # lib/config.py:59

__C.DIR = _AttrDict()
__C.DIR.SHAPENET_QUERY_PATH = './ShapeNet/ShapeNetVox32/'      # lib/config.py:32-36
__C.DIR.MODEL_PATH = './ShapeNet/ShapeNetCore.v1/%s/%s/model.obj'
__C.DIR.VOXEL_PATH = './ShapeNet/ShapeNetVox32/%s/%s/model.binvox'
__C.DIR.RENDERING_PATH = './ShapeNet/ShapeNetRendering/%s/%s/rendering'
__C.DIR.OUT_PATH = './output/default'

__C.TRAIN = _AttrDict()
__C.TRAIN.DATASET_PORTION = [0, 0.8]      # lib/config.py:46-53
__C.TRAIN.NUM_WORKER = 1
__C.TRAIN.NUM_RENDERING = 24
__C.TRAIN.RESUME_TRAIN = False
__C.TRAIN.INITIAL_ITERATION = 0
__C.TRAIN.NUM_ITERATION = 60000
__C.TRAIN.NUM_VALIDATION_ITERATIONS = 24
__C.TRAIN.VALIDATION_FREQ = 2000
__C.TRAIN.NAN_CHECK_FREQ = 2000
__C.TRAIN.RANDOM_NUM_VIEWS = True
__C.TRAIN.DEFAULT_LEARNING_RATE = 1e-4
__C.TRAIN.POLICY = 'adam'
__C.TRAIN.LEARNING_RATES = {'20000': 1e-5, '60000': 1e-6}
__C.TRAIN.MOMENTUM = 0.90
__C.TRAIN.WEIGHT_DECAY = 0.00005
__C.TRAIN.LOSS_LIMIT = 2
__C.TRAIN.SAVE_FREQ = 10000
__C.TRAIN.PRINT_FREQ = 40
# data augmentation (lib/config.py:62-69), used by lib/data_augmentation.py
__C.TRAIN.RANDOM_CROP = True
__C.TRAIN.PAD_X = 10
__C.TRAIN.PAD_Y = 10
__C.TRAIN.FLIP = True
__C.TRAIN.NO_BG_COLOR_RANGE = [[225, 255], [225, 255], [225, 255]]

__C.TEST = _AttrDict()
__C.TEST.DATASET_PORTION = [0.8, 1]        # lib/config.py:94
__C.TEST.EXP_NAME = 'test'                  # lib/config.py:91
__C.TEST.NO_BG_COLOR_RANGE = [[240, 240], [240, 240], [240, 240]]   # lib/config.py:98
__C.TEST.VOXEL_THRESH = [0.4]


def cfg_from_list(cfg_list):
    """Set config keys via list (e.g., from command line): lib/config.py:137-161."""
    assert len(cfg_list) % 2 == 0
    for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
        key_list = k.split('.')
        d = __C
        for subkey in key_list[:-1]:
            assert subkey in d.keys()
            d = d[subkey]
        subkey = key_list[-1]
        assert subkey in d.keys()
        try:
            value = literal_eval(v)
        except Exception:
            value = v
        assert type(value) == type(d[subkey]), \
            'type {} does not match original type {}'.format(type(value), type(d[subkey]))
        d[subkey] = value

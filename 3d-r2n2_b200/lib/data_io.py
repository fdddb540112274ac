"""Dataset index and path templates (lib/data_io.py:15-56)."""
import json
import os
from collections import OrderedDict

from .config import cfg


def id_to_name(id, category_list):
    for k, v in category_list.items():
        if v[0] <= id and v[1] > id:
            return (k, id - v[0])


def category_model_id_pair(dataset_portion=[]):
    """(category id, model id) pairs of the `dataset_portion` = [from, to) fraction of every category listed
    in the cfg.DATASET json, categories and models in sorted order (lib/data_io.py:15-44)."""
    with open(cfg.DATASET) as f:
        cats = OrderedDict(sorted(json.load(f).items(), key=lambda x: x[0]))
    pairs = []
    for _, cat in cats.items():
        model_path = os.path.join(cfg.DIR.SHAPENET_QUERY_PATH, cat['id'])
        models = sorted(name for name in os.listdir(model_path) if os.path.isdir(os.path.join(model_path, name)))
        n = len(models)
        pairs.extend((cat['id'], m) for m in models[int(n * dataset_portion[0]):int(n * dataset_portion[1])])
    print('lib/data_io.py: model paths from %s' % (cfg.DATASET))
    return pairs


def get_model_file(category, model_id):
    return cfg.DIR.MODEL_PATH % (category, model_id)


def get_voxel_file(category, model_id):
    return cfg.DIR.VOXEL_PATH % (category, model_id)


def get_rendering_file(category, model_id, rendering_id):
    return os.path.join(cfg.DIR.RENDERING_PATH % (category, model_id), '%02d.png' % rendering_id)

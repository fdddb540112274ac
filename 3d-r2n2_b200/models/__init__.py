"""Mirror of models/__init__.py (registry + load_model returning the CLASS)."""
from .gru_net import GRUNet
from .res_gru_net import ResidualGRUNet, ResidualLSTMNet

MODELS = (GRUNet, ResidualGRUNet, ResidualLSTMNet)


def get_models():
    '''Returns a tuple of sample models.'''
    return MODELS


def load_model(name):
    '''Returns the model class given its class name (models/__init__.py:12-27).'''
    all_models = get_models()
    mdict = {model.__name__: model for model in all_models}
    if name not in mdict:
        print('Invalid model index. Options are:')
        for model in all_models:
            print('\t* {}'.format(model.__name__))
        return None
    NetClass = mdict[name]
    return NetClass

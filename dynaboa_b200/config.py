"""Asset and dataset paths (same names as reference config.py:7-17; relative to the working directory,
or to ``$DYNABOA_DATA`` when set, so the unchanged driver finds them under ``data/``)."""
import os

_ROOT = os.environ.get('DYNABOA_DATA', 'data')

PW3D_ROOT = os.environ.get('DYNABOA_PW3D_ROOT', 'data/3dpw')
H36M_ROOT = os.environ.get('DYNABOA_H36M_ROOT', 'data/h36m')
InternetData_ROOT = 'supp_assets/bilibili'

DATASET_NPZ_PATH = os.path.join(_ROOT, 'dataset_extras')

JOINT_REGRESSOR_TRAIN_EXTRA = os.path.join(_ROOT, 'J_regressor_extra.npy')
JOINT_REGRESSOR_H36M = os.path.join(_ROOT, 'J_regressor_h36m.npy')
SMPL_MEAN_PARAMS = os.path.join(_ROOT, 'smpl_mean_params.npz')
SMPL_MODEL_DIR = os.path.join(_ROOT, 'smpl')

# files the reference loads with hard-coded paths (base_adaptor.py:55,76-79,116,140)
BASE_MODEL = os.path.join(_ROOT, 'basemodel.pt')
GMM_PRIOR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets', 'gmm_08.npz')
RETRIEVAL_CLUSTERS = os.path.join(_ROOT, 'retrieval_res', 'clusters.pt')
RETRIEVAL_BANK = os.path.join(_ROOT, 'retrieval_res', 'exemplar_bank.pt')


def set_data_root(root):
    """Point every asset path at ``root`` (used by tests / bench with a synthetic asset directory)."""
    global _ROOT, DATASET_NPZ_PATH, JOINT_REGRESSOR_TRAIN_EXTRA, JOINT_REGRESSOR_H36M, SMPL_MEAN_PARAMS, SMPL_MODEL_DIR
    global BASE_MODEL, RETRIEVAL_CLUSTERS, RETRIEVAL_BANK
    _ROOT = root
    DATASET_NPZ_PATH = os.path.join(root, 'dataset_extras')
    JOINT_REGRESSOR_TRAIN_EXTRA = os.path.join(root, 'J_regressor_extra.npy')
    JOINT_REGRESSOR_H36M = os.path.join(root, 'J_regressor_h36m.npy')
    SMPL_MEAN_PARAMS = os.path.join(root, 'smpl_mean_params.npz')
    SMPL_MODEL_DIR = os.path.join(root, 'smpl')
    BASE_MODEL = os.path.join(root, 'basemodel.pt')
    RETRIEVAL_CLUSTERS = os.path.join(root, 'retrieval_res', 'clusters.pt')
    RETRIEVAL_BANK = os.path.join(root, 'retrieval_res', 'exemplar_bank.pt')

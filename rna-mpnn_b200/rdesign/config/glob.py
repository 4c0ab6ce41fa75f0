"""Constants of the rdesign package (same names/values as the reference's
rdesign/config/glob.py:1-32 -- they are part of the numeric and path contract)."""
DATA_PATH: str = './data/'
LOG_PATH: str = './logs/'
OUTPUT_PATH: str = './out/'

FEAT_DIMS: dict = {
    'node': {'angle': 12, 'distance': 80, 'direction': 9},
    'edge': {'orientation': 4, 'distance': 96, 'direction': 15},
}

SPLIT_RATIO: list = [0.9, 0.1, 0.0]
COMPETITION_DATA: str = '/saisdata/'
COMPETITION_OUT: str = '/saisresult/'
BEST_CKPT: str = 'out/checkpoints/RDesign-X/Final-V6.ckpt'
NUM_RES_TYPES: int = 4
DEFAULT_SEED: int = 42
LEPS = 1e6
SEPS = 1e-6

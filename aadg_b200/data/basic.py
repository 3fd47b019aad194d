"""Operator table of the uint8 augmentation bank.

Mirrors the *interface* of the reference's `data/basic.py`: `augment_list()` order defines the op
index the Controller samples (data/basic.py:231-243), `get_augment(name)` gives (id, low, high) and
the level -> value map is `level * (high - low) + low` (data/basic.py:258-260).  There are no
per-image Python functions here: ops are integer ids executed by the CUDA bank
(csrc/aug_u8.cu) on whole batches.
"""

# ids 0..9 are searchable (augment_list order); 10.. are the geometric ops the reference defines but
# never samples (data/basic.py:12-67,82).
AADG_OPS = [
    ("AutoContrast", 0, 1), ("Invert", 0, 1), ("Equalize", 0, 1), ("Solarize", 0, 256),
    ("Posterize", 4, 8), ("Contrast", 0.1, 1.9), ("Color", 0.1, 1.9), ("Brightness", 0.1, 1.9),
    ("Sharpness", 0.1, 1.9), ("Cutout", 0, 0.2),
    ("ShearX", -0.3, 0.3), ("ShearY", -0.3, 0.3), ("TranslateX", -0.45, 0.45),
    ("TranslateY", -0.45, 0.45), ("Rotate", -30, 30), ("Flip", 0, 1),
]
NUM_SEARCHABLE = 10
OP_ID = {name: i for i, (name, _, _) in enumerate(AADG_OPS)}
GEOMETRIC = {"ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate"}
MAX_OPS = 4  # AADG_MAX_OPS in include/aadg_b200.h


def augment_list(for_autoaug=False):
    """[(name, low, high)] of the searchable ops, in the reference's order."""
    if for_autoaug:
        raise NotImplementedError("auto-augment compatibility ops are not part of the hot path")
    return list(AADG_OPS[:NUM_SEARCHABLE])


def get_augment(name):
    """(op id, low, high); KeyError for unknown names (e.g. 'CutMix', like data/basic.py:262-264)."""
    i = OP_ID[name]
    return i, AADG_OPS[i][1], AADG_OPS[i][2]


def level_to_value(name, level):
    _, low, high = get_augment(name)
    return level * (high - low) + low

"""Numeric constants and joint tables of the DynaBOA hot path.

Mirrors the names exported by the reference's ``constants.py`` (reference
constants.py:1-7 camera/normalisation, :15-66 joint names, :70-90 SMPL joint map,
:94-110 selectors and flip permutations) so that ``import constants`` in the
unchanged driver resolves to the same values.  The tables are generated from
compact specs instead of being spelled out entry by entry.
"""

FOCAL_LENGTH = 5000.
IMG_RES = 224
IMG_RES_posenet = 256

IMG_NORM_MEAN = [0.485, 0.456, 0.406]
IMG_NORM_STD = [0.229, 0.224, 0.225]

# (name, index into the 54-joint [24 kinematic | 21 vertex picks | 9 extra-regressor] set)
_OPENPOSE = [
    ('Nose', 24), ('Neck', 12), ('RShoulder', 17), ('RElbow', 19), ('RWrist', 21),
    ('LShoulder', 16), ('LElbow', 18), ('LWrist', 20), ('MidHip', 0), ('RHip', 2),
    ('RKnee', 5), ('RAnkle', 8), ('LHip', 1), ('LKnee', 4), ('LAnkle', 7),
    ('REye', 25), ('LEye', 26), ('REar', 27), ('LEar', 28), ('LBigToe', 29),
    ('LSmallToe', 30), ('LHeel', 31), ('RBigToe', 32), ('RSmallToe', 33), ('RHeel', 34),
]
_GROUND_TRUTH = [
    ('Right Ankle', 8), ('Right Knee', 5), ('Right Hip', 45), ('Left Hip', 46),
    ('Left Knee', 4), ('Left Ankle', 7), ('Right Wrist', 21), ('Right Elbow', 19),
    ('Right Shoulder', 17), ('Left Shoulder', 16), ('Left Elbow', 18), ('Left Wrist', 20),
    ('Neck (LSP)', 47), ('Top of Head (LSP)', 48), ('Pelvis (MPII)', 49),
    ('Thorax (MPII)', 50), ('Spine (H36M)', 51), ('Jaw (H36M)', 52), ('Head (H36M)', 53),
    ('Nose', 24), ('Left Eye', 26), ('Right Eye', 25), ('Left Ear', 28), ('Right Ear', 27),
]

# 25 OpenPose joints followed by the 24 ground-truth joints (49 total).
JOINT_NAMES = ['OP ' + n for n, _ in _OPENPOSE] + [n for n, _ in _GROUND_TRUTH]
JOINT_IDS = {name: i for i, name in enumerate(JOINT_NAMES)}
JOINT_MAP = {**{'OP ' + n: j for n, j in _OPENPOSE}, **{n: j for n, j in _GROUND_TRUTH}}

# The 49 source indices in output order (what ``SMPL.joint_map`` holds).
JOINT_MAP_49 = [JOINT_MAP[n] for n in JOINT_NAMES]

# 14 LSP joints out of the 17 H36M joints / out of the 24 ground-truth joints.
H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]
H36M_TO_J14 = H36M_TO_J17[:14]
J24_TO_J17 = list(range(13)) + [18, 14, 16, 17]
J24_TO_J14 = J24_TO_J17[:14]

# Left/right flip permutations.
SMPL_JOINTS_FLIP_PERM = [0, 2, 1, 3, 5, 4, 6, 8, 7, 9, 11, 10, 12, 14, 13, 15, 17, 16,
                         19, 18, 21, 20, 23, 22]
SMPL_POSE_FLIP_PERM = [3 * j + c for j in SMPL_JOINTS_FLIP_PERM for c in range(3)]
J24_FLIP_PERM = [5, 4, 3, 2, 1, 0, 11, 10, 9, 8, 7, 6, 12, 13, 14, 15, 16, 17, 18, 19,
                 21, 20, 23, 22]
J49_FLIP_PERM = [0, 1, 5, 6, 7, 2, 3, 4, 8, 12, 13, 14, 9, 10, 11, 16, 15, 18, 17, 22, 23,
                 24, 19, 20, 21] + [25 + i for i in J24_FLIP_PERM]

# SMPL kinematic tree (smplx model data; 24 joints, parent of the root is -1).
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

# smplx ``vertex_ids['smplh']`` picks appended after the 24 kinematic joints
# (smplx/vertex_joint_selector.py order: face, feet, then finger tips l/r).
SMPL_EXTRA_VERTEX_IDS = [
    332, 6260, 2800, 4071, 583,              # nose, reye, leye, rear, lear
    3216, 3226, 3387, 6617, 6624, 6787,      # LBigToe, LSmallToe, LHeel, RBigToe, RSmallToe, RHeel
    2746, 2319, 2445, 2556, 2673,            # left thumb, index, middle, ring, pinky
    6191, 5782, 5905, 6016, 6133,            # right thumb, index, middle, ring, pinky
]

NUM_VERTS = 6890

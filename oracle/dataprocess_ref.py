"""Oracle restatement of the input side of the hot path (test infrastructure only): reference utils/dataprocess.py
``get_transform`` :13-37, ``transform`` :39-46, ``crop`` :48-96 (rot = 0: the benchmark path uses no augmentation,
boa_dataset/pw3d.py:103-106) and boa_dataset/pw3d.py ``rgb_processing`` :144-149, ``j2d_processing`` :151-163,
``process_sample`` :127-136 (torchvision Normalize with constants.IMG_NORM_MEAN / STD).

``skimage.transform.resize`` (unpinned dependency, not installed here) is restated through the two scipy.ndimage calls it is
made of in skimage >= 0.19 for order 1 without rotation (skimage/transform/_warps.py::resize): a Gaussian pre-filter with
sigma = max(0, (in / out - 1) / 2) per axis, mode 'reflect' -> ndimage 'mirror', then ``ndi.zoom(..., order=1, mode='mirror',
grid_mode=True)``.  ``oracle/make_golden.py dataprocess`` runs the REFERENCE's own crop / transform code with only that
resize stubbed by ``resize`` below and asserts this module reproduces it.
"""
import numpy as np
import scipy.ndimage as ndi

IMG_RES = 224
IMG_NORM_MEAN = np.array([0.485, 0.456, 0.406])
IMG_NORM_STD = np.array([0.229, 0.224, 0.225])


def get_transform(center, scale, res):
    h = 200 * scale
    t = np.zeros((3, 3))
    t[0, 0] = float(res[1]) / h
    t[1, 1] = float(res[0]) / h
    t[0, 2] = res[1] * (-float(center[0]) / h + .5)
    t[1, 2] = res[0] * (-float(center[1]) / h + .5)
    t[2, 2] = 1
    return t


def transform(pt, center, scale, res, invert=0):
    t = get_transform(center, scale, res)
    if invert:
        t = np.linalg.inv(t)
    new_pt = np.dot(t, np.array([pt[0] - 1, pt[1] - 1, 1.]).T)
    return new_pt[:2].astype(int) + 1


def resize(image, res):
    """skimage.transform.resize(image, res) with its defaults (order 1, mode 'reflect', anti_aliasing on down-scaling)."""
    image = np.asarray(image, dtype=np.float64)
    factors = np.array(image.shape[:2], dtype=np.float64) / np.array(res, dtype=np.float64)
    sigma = np.maximum(0, (factors - 1) / 2)
    if image.ndim == 3:
        sigma = np.append(sigma, 0.0)
    filtered = ndi.gaussian_filter(image, sigma, cval=0, mode='mirror') if np.any(sigma > 0) else image
    zoom = [1.0 / factors[0], 1.0 / factors[1]] + ([1.0] if image.ndim == 3 else [])
    return ndi.zoom(filtered, zoom, order=1, mode='mirror', cval=0, grid_mode=True)


def crop_box(center, scale, res):
    ul = np.array(transform([1, 1], center, scale, res, invert=1)) - 1
    br = np.array(transform([res[0] + 1, res[1] + 1], center, scale, res, invert=1)) - 1
    return ul, br


def crop(img, center, scale, res):
    ul, br = crop_box(center, scale, res)
    new_img = np.zeros([br[1] - ul[1], br[0] - ul[0], img.shape[2]])
    new_x = max(0, -ul[0]), min(br[0], len(img[0])) - ul[0]
    new_y = max(0, -ul[1]), min(br[1], len(img)) - ul[1]
    old_x = max(0, ul[0]), min(len(img[0]), br[0])
    old_y = max(0, ul[1]), min(len(img), br[1])
    new_img[new_y[0]:new_y[1], new_x[0]:new_x[1]] = img[old_y[0]:old_y[1], old_x[0]:old_x[1]]
    return resize(new_img, res)


def rgb_processing(rgb_img, center, scale):
    """(H, W, 3) float32 RGB 0..255 -> normalised (3, 224, 224) float32 (pw3d.py:144-149 + Normalize :133)."""
    img = crop(rgb_img.copy(), center, scale, [IMG_RES, IMG_RES])
    img = np.transpose(img.astype('float32'), (2, 0, 1)) / 255.0
    return ((img - IMG_NORM_MEAN[:, None, None].astype('float32')) / IMG_NORM_STD[:, None, None].astype('float32')).astype('float32')


def j2d_processing(kp, center, scale):
    """(N, 3) pixel keypoints (+ confidence) -> integer crop pixels mapped to [-1, 1] (pw3d.py:151-163)."""
    kp = kp.copy()
    for i in range(kp.shape[0]):
        kp[i, 0:2] = transform(kp[i, 0:2] + 1, center, scale, [IMG_RES, IMG_RES])
    kp[:, :-1] = 2. * kp[:, :-1] / IMG_RES - 1.
    return kp.astype('float32')

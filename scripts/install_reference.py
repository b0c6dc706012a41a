"""Copy the UNMODIFIED reference sources the GPU box needs into the git-ignored ``baseline/_ref/`` (it travels with gpurun;
``/root/reference`` does not exist there).  Nothing here is product source: ``baseline/_ref/dynaboa_benchmark.py`` is
executed as-is on top of the drop-in module tree by tests/test_gpu_dropin_driver.py, and the remaining files are the
reference's own CPU implementation of the hot path timed by ``bench.py --impl reference`` (oracle/ref_harness.py).

Run in the build container:  python scripts/install_reference.py   (also called by __graft_entry__.build()).
"""
import hashlib
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
DST = os.path.join(REPO, 'baseline', '_ref')
FILES = ['dynaboa_benchmark.py', 'dynaboa_internet.py', 'base_adaptor.py', 'constants.py', 'config.py', 'model/__init__.py', 'model/hmr.py',
         'model/smpl.py', 'utils/__init__.py', 'utils/geometry.py', 'utils/pose_utils.py', 'utils/dataprocess.py', 'utils/smplify/__init__.py',
         'utils/smplify/prior.py', 'utils/smplify/smplify.py', 'utils/smplify/losses.py', 'boa_dataset/__init__.py', 'boa_dataset/pw3d.py',
         'boa_dataset/internet_data.py', 'data/gmm_08.pkl']


def install():
    if not os.path.isdir(REF):
        return False
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append(f'{hashlib.sha256(open(src, "rb").read()).hexdigest()}  {rel}')
    with open(os.path.join(DST, 'MANIFEST.sha256'), 'w') as f:
        f.write('\n'.join(manifest) + '\n')
    return True


if __name__ == '__main__':
    ok = install()
    print('installed' if ok else 'no /root/reference here', DST)
    sys.exit(0)

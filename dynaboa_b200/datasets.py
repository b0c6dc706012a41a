"""Frame sources with the dict keys of the reference datasets (boa_dataset/pw3d.py:115-125,
boa_dataset/internet_data.py).  Image decoding / cropping of real 3DPW or Internet videos is the step BEFORE
the hot path and is out of scope here; these classes serve the seeded synthetic stream of
``dynaboa_b200.synthetic`` so that the drivers run end to end without the licensed datasets."""
from torch.utils.data import Dataset

from . import synthetic


class PW3D(Dataset):
    def __init__(self, options=None):
        n = getattr(options, 'synthetic_frames', 16) if options is not None else 16
        seed = getattr(options, 'seq_seed', synthetic.SEED) if options is not None else synthetic.SEED
        rank = getattr(options, 'rank', 0) if options is not None else 0
        self.stream = synthetic.SyntheticStream(length=n, batch_size=1, seed=seed, rank=rank)

    def __len__(self):
        return len(self.stream)

    def __getitem__(self, i):
        f = self.stream[i]
        item = {}
        for k, v in f.items():
            item[k] = v[0]
        return item


class Internet_dataset(PW3D):
    """Same stream without ground-truth evaluation fields being meaningful (reference internet driver)."""

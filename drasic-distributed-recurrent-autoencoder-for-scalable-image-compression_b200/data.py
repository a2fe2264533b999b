"""Stand-in for the reference's data.py (fetch_dataset / make_data_loader, data.py:13-83) so its training and test
loops run end to end against this package.  The reference's `datasets` package (torchvision
wrappers with downloaders and an anytree class hierarchy) is outside the coding loop and not rebuilt; what
the loop consumes is a stream of {'img': [C,32,32] in [0,1], 'label': node id} samples, produced here
deterministically.  `subset` keeps the reference's meaning of "which annotation becomes the node id"."""
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.dataloader import default_collate

import config

DEFAULT_SIZE = {'train': 2048, 'test': 512}
CHANNELS = {'MNIST': 1, 'FashionMNIST': 1, 'EMNIST': 1, 'SVHN': 3, 'CIFAR10': 3, 'CIFAR100': 3, 'Kodak': 3}


class SyntheticImages(Dataset):
    """smooth random images (low-frequency noise upsampled, so a codec has something to learn) with a node id
    per sample; sample i is a pure function of (split, i)"""

    def __init__(self, data_name, split, size, num_node, subset='label'):
        if data_name not in CHANNELS:
            raise ValueError('Not valid dataset name')
        self.data_name, self.split, self.size, self.subset = data_name, split, int(size), subset
        self.channels, self.num_node = CHANNELS[data_name], max(1, int(num_node))
        self.transform = None
        self._salt = {'train': 0, 'test': 1 << 20}[split]

    def __len__(self):
        return self.size

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self._salt + int(index))
        coarse = torch.rand(1, self.channels, 4, 4, generator=g)
        img = torch.nn.functional.interpolate(coarse, size=(32, 32), mode='bilinear', align_corners=False)[0]
        img = (img + 0.05 * torch.rand(self.channels, 32, 32, generator=g)).clamp_(0, 1)
        label = torch.tensor(int(torch.randint(0, self.num_node, (1,), generator=g)))
        return {'img': img, 'label': label}


def fetch_dataset(data_name, subset):
    num_node = int(config.PARAM['control']['num_node'])
    sizes = {k: (config.PARAM['data_size'][k] or DEFAULT_SIZE[k]) for k in ('train', 'test')}
    return {k: SyntheticImages(data_name, k, sizes[k], num_node, subset) for k in ('train', 'test')}


def input_collate(batch):
    """dict samples -> dict of lists (utils.collate stacks them, train_model.py:109); anything else -> default"""
    if isinstance(batch[0], dict):
        return {key: [b[key] for b in batch] for key in batch[0]}
    return default_collate(batch)


def make_data_loader(dataset):
    return {k: DataLoader(dataset=dataset[k], shuffle=config.PARAM['shuffle'][k], batch_size=config.PARAM['batch_size'][k],
                          pin_memory=True, num_workers=config.PARAM['num_workers'], collate_fn=input_collate)
            for k in dataset}

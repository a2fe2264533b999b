"""Generate the golden fixtures in this directory by running the REFERENCE's own modules.

Run only in the build container (the reference tree is mounted at /root/reference; it does not exist
on the GPU box):      python tests/golden/make_golden.py
Follows the oracle recipe of SURVEY.md 8c: import modules/functions/models.dist_codec from
/root/reference/src unmodified, stub the missing `models.magick`, and drive the cells with an outer
loop that restates models/dist_codec.py:36-65 verbatim except line 52 (self-aliasing index_put that
raises on torch >= 1.5) which becomes an out-of-place scatter.
Weights are NOT stored (81 M parameters at M=8): they are the reference constructors' default
initialisation under torch.manual_seed(seed); each fixture stores a sha256 over the state_dict so a
consumer can prove it regenerated the same tensors.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/src'
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, REF)
    import config
    config.init()
    config.PARAM['device'] = 'cpu'
    sys.modules['models.magick'] = types.ModuleType('models.magick')
    import models      # noqa: F401
    import modules     # noqa: F401
    import functions   # noqa: F401
    import utils       # noqa: F401
    return config, models, modules, functions, utils


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_model(config, models, utils, name, data_name, num_node, num_iter, code_size=8, seed=0):
    config.PARAM['data_name'] = data_name
    config.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': str(num_iter),
                               'code_size': str(code_size), 'num_node': str(num_node)}
    utils.process_control_name()
    torch.manual_seed(seed)
    cwd = os.getcwd()
    os.chdir('/tmp')  # dist_codec.Model.__init__ creates ./output/tmp
    try:
        m = getattr(models, name)()
    finally:
        os.chdir(cwd)
    return m


def drive_dist(config, m, img, label, training, noise_seed=None):
    """models/dist_codec.py:36-65 with line 52 made out-of-place; records z, codes, noise."""
    m.train(training)
    zs, codes, noises = [], [], []

    def hook(mod, inp, out):
        zs.append(inp[0].detach().clone())
        codes.append(out.detach().clone())
    hdl = m.model['quantizer'].register_forward_hook(hook)
    output = {'loss': torch.tensor(0, device='cpu', dtype=torch.float32), 'img': [], 'code': []}
    x = img * 2 - 1
    reconstructed = x.new_zeros(x.size())
    indices = torch.arange(img.size(0))
    T, M = config.PARAM['num_iter'], config.PARAM['num_node']
    for i in range(T):
        encoded, node_indices = [], []
        for j in range(M):
            node_x = x[label == j]
            if node_x.size(0) == 0:
                continue
            encoded.append(m.model['encoder'][j](node_x))
            node_indices.append(indices[label == j])
        encoded = torch.cat(encoded, dim=0)
        node_indices = torch.cat(node_indices, dim=0)
        tmp = torch.empty_like(encoded)
        tmp[node_indices] = encoded
        encoded = tmp
        if training:
            # the reference draws `x.new(x.size()).uniform_()` inside Quantize; reseed so the same
            # uniforms can be regenerated and stored
            torch.manual_seed(noise_seed + i)
            noises.append(torch.empty_like(encoded).uniform_())
            torch.manual_seed(noise_seed + i)
        code = m.model['quantizer'](encoded)
        output['code'].append(m.pack(code))
        decoded = m.model['decoder'](code)
        decoded = (decoded + 1) / 2 if i == 0 else decoded
        reconstructed = reconstructed + decoded
        output['loss'] = output['loss'] + m.loss_fn(reconstructed, img)
        x = img - reconstructed
        output['img'].append(reconstructed)
    output['loss'] = output['loss'] / T
    for i in range(1, T):
        output['code'][i] = np.concatenate((output['code'][i - 1], output['code'][i]), axis=1)
    hdl.remove()
    return output, zs, codes, noises


def drive_sep(config, m, img, label):
    """models/sep_codec.py:35-66 (eval), line 57 made out-of-place."""
    m.train(False)
    output = {'loss': torch.tensor(0, device='cpu', dtype=torch.float32), 'img': [], 'code': []}
    x = img * 2 - 1
    reconstructed = x.new_zeros(x.size())
    indices = torch.arange(img.size(0))
    T, M = config.PARAM['num_iter'], config.PARAM['num_node']
    all_codes = []
    for i in range(T):
        code, decoded, node_indices, raw = [], [], [], []
        for j in range(M):
            node_x = x[label == j]
            if node_x.size(0) == 0:
                continue
            encoded = m.model['encoder'][j](node_x)
            e_code = m.model['quantizer'](encoded)
            raw.append(e_code.detach())
            code.append(m.pack(e_code))
            decoded.append(m.model['decoder'][j](e_code))
            decoded[-1] = (decoded[-1] + 1) / 2 if i == 0 else decoded[-1]
            node_indices.append(indices[label == j])
        output['code'].append(np.concatenate(code, axis=0))
        decoded = torch.cat(decoded, dim=0)
        node_indices = torch.cat(node_indices, dim=0)
        tmp = torch.empty_like(decoded)
        tmp[node_indices] = decoded
        decoded = tmp
        rawc = torch.cat(raw, 0)
        tmpc = torch.empty_like(rawc)
        tmpc[node_indices] = rawc
        all_codes.append(tmpc)
        reconstructed = reconstructed + decoded
        output['loss'] = output['loss'] + m.loss_fn(reconstructed, img)
        x = img - reconstructed.detach()
        output['img'].append(reconstructed.detach())
    output['loss'] = output['loss'] / T
    return output, all_codes


def psnr_list(metrics_psnr, imgs, img):
    return np.array([metrics_psnr(r.detach(), img) for r in imgs], dtype=np.float64)


def main():
    torch.set_num_threads(1)  # single-thread reference results are the pinned ones
    config, models, modules, functions, utils = import_reference()
    sys.modules.setdefault('metrics_ref', types.ModuleType('metrics_ref'))
    import importlib.util
    spec = importlib.util.spec_from_file_location('metrics_ref', os.path.join(REF, 'metrics.py'))
    metrics = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(metrics)

    # ------------------------------------------------ 1. functions + single cells
    g = torch.Generator().manual_seed(11)
    xs = torch.rand(2, 8, 6, 4, generator=g)
    fx = {
        'shuffle_in': xs.numpy(),
        'unshuffle_out': functions.pixel_unshuffle(xs, 2).numpy(),
        'shuffle_out': functions.pixel_shuffle(xs, 2).numpy(),
    }
    q_in = torch.tensor([-2.0, -1.0, -0.5, -0.0, 0.0, 1e-8, -1e-8, 0.5, 1.0, 1.5, float('nan')])
    fx['quant_in'] = q_in.numpy()
    fx['quant_eval'] = functions.Quantize.apply(q_in, False).numpy()
    qt = torch.rand(64, generator=g) * 2 - 1
    torch.manual_seed(5)
    u = torch.empty_like(qt).uniform_()
    torch.manual_seed(5)
    fx['quant_train_in'] = qt.numpy()
    fx['quant_train_u'] = u.numpy()
    fx['quant_train'] = functions.Quantize.apply(qt, True).numpy()

    for tag, (cin, co, hw) in {'lstm_a': (128, 128, 4), 'lstm_b': (64, 64, 8)}.items():
        torch.manual_seed(21)
        cell = modules.Cell({'cell': 'ConvLSTMCell', 'input_size': cin, 'output_size': co, 'kernel_size': 3,
                             'stride': 1, 'padding': 1, 'bias': False, 'num_layers': 1, 'activation': 'tanh'})
        gi = torch.Generator().manual_seed(22)
        hs, cs = [], []
        with torch.no_grad():
            for t in range(3):
                x = torch.rand(2, cin, hw, hw, generator=gi) * 2 - 1
                h = cell(x)
                hs.append(h.numpy().copy())
                cs.append(cell.cell.hidden[1][0].numpy().copy())
        fx[tag + '_h'] = np.stack(hs)
        fx[tag + '_c'] = np.stack(cs)
        fx[tag + '_digest'] = np.array(sd_digest(cell.state_dict()))
    torch.manual_seed(31)
    cc = modules.Cell({'cell': 'ConvCell', 'input_size': 3, 'output_size': 32, 'kernel_size': 1, 'stride': 1,
                       'padding': 0, 'bias': True, 'activation': 'tanh', 'normalization': 'none'})
    xc = torch.rand(2, 3, 8, 8, generator=g)
    with torch.no_grad():
        fx['convcell_in'] = xc.numpy()
        fx['convcell_out'] = cc(xc).numpy()
        fx['convcell_w'] = cc.state_dict()['cell.cell.main.weight'].numpy()
        fx['convcell_b'] = cc.state_dict()['cell.cell.main.bias'].numpy()
    np.savez_compressed(os.path.join(OUT, 'cells.npz'), **fx)

    # ------------------------------------------------ 2. config 1: MNIST-like, single source, eval
    m = build_model(config, models, utils, 'dist_codec', 'MNIST', 1, 16, seed=0)
    gi = torch.Generator().manual_seed(0)
    img = torch.rand(4, 1, 32, 32, generator=gi)
    label = torch.zeros(4, dtype=torch.long)
    with torch.no_grad():
        out, zs, codes, _ = drive_dist(config, m, img, label, False)
    np.savez_compressed(
        os.path.join(OUT, 'loop_mnist_m1.npz'),
        digest=np.array(sd_digest(m.state_dict())), keys=np.array(list(m.state_dict().keys())),
        img=img.numpy(), label=label.numpy(),
        codes=torch.stack(codes).numpy().astype(np.int8), z=torch.stack(zs).numpy(),
        recon_first=out['img'][0].numpy(), recon_last=out['img'][-1].numpy(),
        psnr=psnr_list(metrics.PSNR, out['img'], img), loss=np.float32(out['loss'].item()),
        code_nbytes=np.array([c.nbytes for c in out['code']]), code_shape_last=np.array(out['code'][-1].shape),
        code_byte_values=np.unique(out['code'][-1]))

    # ------------------------------------------------ 3. CIFAR-like, 2 sources, eval + train (+grads)
    m = build_model(config, models, utils, 'dist_codec', 'CIFAR10', 2, 4, seed=3)
    gi = torch.Generator().manual_seed(4)
    img = torch.rand(6, 3, 32, 32, generator=gi)
    label = torch.tensor([0, 1, 1, 0, 1, 0])
    with torch.no_grad():
        out, zs, codes, _ = drive_dist(config, m, img, label, False)
    fx = dict(digest=np.array(sd_digest(m.state_dict())), img=img.numpy(), label=label.numpy(),
              codes=torch.stack(codes).numpy().astype(np.int8), z=torch.stack(zs).numpy(),
              recon=torch.stack([r for r in out['img']]).numpy(),
              psnr=psnr_list(metrics.PSNR, out['img'], img), loss=np.float32(out['loss'].item()))
    import utils as ref_utils
    ref_utils.apply_fn(m, 'free_hidden')
    m.zero_grad()
    out, zs, codes, noises = drive_dist(config, m, img, label, True, noise_seed=100)
    out['loss'].backward()
    fx.update(train_noise=torch.stack(noises).numpy(), train_codes=torch.stack(codes).numpy().astype(np.int8),
              train_loss=np.float32(out['loss'].item()),
              train_recon_last=out['img'][-1].detach().numpy())
    gn = {}
    for k, p_ in m.named_parameters():
        gn[k] = float(p_.grad.double().norm())
    fx['grad_keys'] = np.array(list(gn.keys()))
    fx['grad_norms'] = np.array(list(gn.values()))
    # a few full gradients (small tensors) for element-wise comparison
    named = dict(m.named_parameters())
    for short, k in {'g_e0_w': 'model.encoder.0.0.cell.cell.main.weight',
                     'g_e0_b': 'model.encoder.1.0.cell.cell.main.bias',
                     'g_e4_w': 'model.encoder.1.7.cell.cell.main.weight',
                     'g_d0_w': 'model.decoder.0.cell.cell.main.weight',
                     'g_d4_w': 'model.decoder.7.cell.cell.main.weight'}.items():
        fx[short] = named[k].grad.numpy()
    # one slice of a big ConvLSTM gradient
    fx['g_d3_in_slice'] = named['model.decoder.5.cell.cell.0.in.cell.cell.main.weight'].grad[:8, :8].numpy()
    fx['g_e1_hidden_slice'] = named['model.encoder.0.2.cell.cell.0.hidden.cell.cell.main.weight'].grad[:8, :8].numpy()
    ref_utils.apply_fn(m, 'free_hidden')
    np.savez_compressed(os.path.join(OUT, 'loop_cifar_m2.npz'), **fx)

    # ------------------------------------------------ 4. sep_codec, 2 sources, eval
    m = build_model(config, models, utils, 'sep_codec', 'CIFAR10', 2, 3, seed=6)
    gi = torch.Generator().manual_seed(7)
    img = torch.rand(5, 3, 32, 32, generator=gi)
    label = torch.tensor([1, 0, 1, 1, 0])
    with torch.no_grad():
        out, codes = drive_sep(config, m, img, label)
    np.savez_compressed(
        os.path.join(OUT, 'loop_sep_m2.npz'),
        digest=np.array(sd_digest(m.state_dict())), keys=np.array(list(m.state_dict().keys())),
        img=img.numpy(), label=label.numpy(),
        codes=torch.stack(codes).numpy().astype(np.int8), recon=torch.stack(out['img']).numpy(),
        psnr=psnr_list(metrics.PSNR, out['img'], img), loss=np.float32(out['loss'].item()))
    print('golden fixtures written to', OUT)


if __name__ == '__main__':
    main()

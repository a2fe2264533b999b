"""Global experiment dictionary, same keys and defaults as the reference's src/config.py:1-33 (the
caller scripts turn every key into a command-line flag, train_model.py:20-36).  Two optional keys are
B200-specific: 'precision' ('parity' = fp32-grade split-fp16 tensor-core path with bit-exact codes,
'fast' = single fp16) and 'pack_mode' ('reference' = the reference's np.packbits semantics, 'bits' =
the real 1 bit/symbol stream), 'grad_precision' ('fp16' | 'parity') and 'cuda_graph' (capture the whole T-iteration
loop, forward and backward, as one CUDA graph per batch shape).  Further optional keys read with .get(): 'cta_pair',
'stream_k' ('auto' | 'on' | 'off'), 'node_iters' (per-node iteration budgets) and 'deterministic' (True: gradients are
bit-identical from run to run, like the reference's CPU path; a few per cent slower)."""


def init():
    global PARAM
    import compat
    compat.apply()
    PARAM = dict(
        data_name='MNIST', subset='label', model_name='dist_codec',
        control=dict(normalization='none', activation='tanh', num_iter='16', code_size='8', num_node='10'),
        optimizer_name='Adam', lr=1e-3, momentum=0, weight_decay=0,
        scheduler_name='ReduceLROnPlateau', step_size=1, milestones=[150, 250], patience=5, threshold=1e-3,
        factor=0.1, batch_size=dict(train=10, test=200), shuffle=dict(train=True, test=False), num_workers=0,
        data_size=dict(train=0, test=0), device='cuda', num_epochs=200, save_mode=0, world_size=1,
        metric_names=dict(train=['Loss', 'BPP', 'PSNR'], test=['Loss', 'BPP', 'PSNR']),
        init_seed=0, num_Experiments=1, log_interval=0.25, log_overwrite=False, loss_fn='mae', resume_mode=0,
        precision='parity', grad_precision='fp16', pack_mode='reference', cuda_graph=False)

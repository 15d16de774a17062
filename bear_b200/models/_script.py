"""Shared body of the config-driven scripts train_bear_net.py / train_bear_ref.py.

Keeps the reference scripts' contract (models/train_bear_net.py:29-200, models/train_bear_ref.py):
same INI sections and keys, the magic values ``out_folder = TEST`` / trailing ``*`` and
``files_path = TEST``, the model-file layout ``<out>/config.cfg`` (input + ``[results]``) and
``<out>/results.pickle`` = ``dill.dump({'params': [h_signed, *ar_params]})``, and the return value
``1`` or ``(1, ll_van, perp_van)``.  Parameters are stored as numpy arrays (objects only need to be
array-like for ``change_scope_params``).
"""
import datetime
import json
import os
import sys

import numpy as np
import torch

from .. import ar_funcs, core, dataloader

HERE = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(os.path.dirname(HERE), 'data')


def _out_folder(config):
    stamp = datetime.datetime.now().strftime("%Y%m%d-%H%M%S")
    of = config['general']['out_folder']
    if of == 'TEST':
        return os.path.join(HERE, 'out_data', 'logs', stamp)
    if of.endswith('*'):
        return of[:-1]
    return os.path.join(of, 'logs', stamp)


def _writer(out_folder):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(out_folder)
    except Exception:
        return None


def _write_config(config, out_folder):
    with open(os.path.join(out_folder, 'config.cfg'), 'w') as cw:
        config.write(cw)


def _save_loss_plot(loss_save, out_folder):
    try:
        import matplotlib
        matplotlib.use('Agg')
        from matplotlib import pyplot as plt
    except Exception:
        np.savetxt(os.path.join(out_folder, 'loss.txt'), np.asarray(loss_save))
        return
    plt.figure(figsize=[10, 10])
    plt.xlabel("steps", fontsize=30)
    plt.ylabel("loss", fontsize=30)
    plt.plot(loss_save)
    plt.tight_layout()
    plt.savefig(os.path.join(out_folder, 'loss.png'), dpi=200)
    plt.close()


def run(config, model, is_ref):
    """``model`` is the bear_net or bear_ref module."""
    import dill
    out_folder = _out_folder(config)
    os.makedirs(out_folder, exist_ok=True)
    torch.manual_seed(int(config['general']['seed']))
    precision = config['general']['precision']
    if precision not in ('float64', 'float32'):
        raise ValueError("[general] precision must be float64 or float32, not '%s'" % precision)
    if precision == 'float32':
        # models/train_bear_net.py:43 lets the TF graph run in float32.  Every kernel here computes in float64, a
        # superset: a float32 config runs unchanged and its results agree with a float32 reference run to float32
        # rounding (~1e-6 relative on log-likelihoods); there is no reduced-precision path to select.
        print('precision = float32 requested: computing in float64 (results agree to float32 rounding)', file=sys.stderr)
    writer = _writer(out_folder)

    # Load data.
    if config['data']['files_path'] == 'TEST':
        files = [os.path.join(DATA_DIR, 'ysd1_lag_5_file_0_preshuf.tsv')]
    else:
        fp = config['data']['files_path']
        files = [os.path.join(fp, f) for f in os.listdir(fp) if f.startswith(config['data']['start_token'])]
    sparse = config['data']['sparse'] == 'True'
    num_kmers = sum(dataloader.count_rows(f, header=sparse) for f in files)
    kmer_batch_size = float(config['train']['batch_size'])
    kmer_batch_size = int(num_kmers * kmer_batch_size) if kmer_batch_size <= 1 else int(kmer_batch_size)
    epochs = config['train']['epochs']
    if epochs[-1] == 's':
        epochs = int(epochs[:-1]) // (1 + num_kmers // kmer_batch_size) + 1
    else:
        epochs = int(epochs)
    num_ds = int(config['data']['num_ds'])
    alphabet = config['data']['alphabet']
    data = dataloader.load_files(files, alphabet, kmer_batch_size, num_ds, sparse=sparse)
    data_train = data.repeat(epochs)

    result_file = os.path.join(out_folder, 'results.pickle')
    config['results']['out_folder'] = out_folder
    config['results']['file'] = result_file
    _write_config(config, out_folder)

    ds_loc = int(config['data']['train_column'])
    ds_loc_ref = int(config['data']['reference_column']) if is_ref else None
    alphabet_size = len(core.alphabets_tf[alphabet]) - 1
    lag = int(config['hyperp']['lag'])
    make_ar_func = getattr(ar_funcs, 'make_ar_func_' + config['model']['ar_func_name'])
    af_kwargs = json.loads(config['model']['af_kwargs'])
    learning_rate = float(config['train']['learning_rate'])
    optimizer_name = config['train']['optimizer_name']
    train_ar = config['train']['train_ar'] == 'True'
    acc_steps = int(config['train']['accumulation_steps'])

    if config['train']['restart'] == 'True':
        with open(os.path.join(config['train']['restart_path'], "results.pickle"), 'rb') as fr:
            params_restart = dill.load(fr)['params']
    else:
        params_restart = None
    ref_args = (ds_loc_ref,) if is_ref else ()

    if config['train']['train'] == 'True':
        loss_save = []
        params, h_signed, ar_func = model.train(
            data_train, num_kmers, epochs, ds_loc, *ref_args, alphabet, lag, make_ar_func, af_kwargs,
            learning_rate, optimizer_name, train_ar=train_ar, acc_steps=acc_steps,
            params_restart=params_restart, writer=writer, loss_save=loss_save)
        _save_loss_plot(loss_save, out_folder)
    else:
        assert config['train']['restart'] == 'True'
        params, h_signed, ar_func = model.change_scope_params(lag, alphabet_size, make_ar_func, af_kwargs, params_restart)

    h = torch.exp(h_signed)
    config['results']['h'] = str(float(h))
    if is_ref:
        tau = torch.exp(params[1])
        config['results']['error_rate'] = str(float(1 - torch.exp(-tau)))
        nw = torch.exp(params[2])
        config['results']['stop_rate'] = str(float(1 / (nw / (1 + nw))))
    _write_config(config, out_folder)
    with open(result_file, 'wb') as rw:
        dill.dump({'params': [p.detach().cpu().numpy() for p in params]}, rw)

    def record(prefix, out):
        ll_ear, ll_ar, ll_van, perp_ear, perp_ar, perp_van, acc_ear, acc_ar, acc_van = out
        r = config['results']
        r[prefix + 'perplex_BEAR'] = str(float(perp_ear))
        r[prefix + 'perplex_AR'] = str(float(perp_ar))
        r[prefix + 'perplex_BMM'] = json.dumps(perp_van.numpy().tolist())
        r[prefix + 'loglikelihood_BEAR'] = str(float(ll_ear))
        r[prefix + 'loglikelihood_AR'] = str(float(ll_ar))
        r[prefix + 'loglikelihood_BMM'] = json.dumps(ll_van.numpy().tolist())
        r[prefix + 'accuracy_BEAR'] = str(float(acc_ear))
        r[prefix + 'accuracy_AR'] = str(float(acc_ar))
        r[prefix + 'accuracy_BMM'] = json.dumps(acc_van.numpy().tolist())
        _write_config(config, out_folder)

    if config['test']['test'] == 'True':
        ds_loc_test = int(config['data']['test_column'])
        van_reg = np.array(json.loads(config['test']['van_reg']))
        record('heldout_', model.evaluation(data, ds_loc, ds_loc_test, *ref_args, alphabet, h, ar_func, van_reg))

    if config['test']['train_test'] == 'True':
        van_reg = np.array(json.loads(config['test']['van_reg']))
        out = model.evaluation(data, -1, ds_loc, *ref_args, alphabet, h, ar_func, van_reg)
        record('', out)
        return 1, out[2].numpy(), out[5].numpy()
    return 1

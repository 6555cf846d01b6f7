"""Entry point with the reference's command line (reference src/main.py).

    python main.py --model_name BPRMF    --emb_size 64 --lr 1e-3 --l2 1e-6 --dataset ml-100k
    python main.py --model_name LightGCN --emb_size 64 --gcn_layers 2 --lr 2e-3 --dataset ml-100k

Same two-stage argument parsing (unknown flags are ignored, main.py:109,122 -- so `--emb_size`, which the
general models never defined, is silently dropped here too; the working flag is `--embedding_size`), same
log / checkpoint naming, same flow: seed -> reader -> model -> datasets -> runner.train -> test.  The one
behavioural difference is the device: there is no CPU path, a CUDA device is required.
"""
import argparse
import importlib
import logging
import os
import pickle
import sys

if __package__ in (None, ''):                      # `python main.py` from inside the package directory
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    __package__ = 'whisprrec_b200'

import torch

from .utils import utils

MODEL_PACKAGES = ('models.general',)
HELPER_PACKAGE = 'helpers'


def resolve(kind, name):
    """The reference resolves `Name.Name` from star-imports (main.py:110-114): module and class share a name."""
    packages = MODEL_PACKAGES if kind == 'model' else (HELPER_PACKAGE,)
    for pkg in packages:
        try:
            module = importlib.import_module('.{}.{}'.format(pkg, name), package=__package__)
        except ModuleNotFoundError:
            continue
        return getattr(module, name)
    raise NameError("name '{}' is not defined".format(name))


class _CorpusUnpickler(pickle.Unpickler):
    """Corpus caches written by the reference (`../data/<dataset>/BaseReader.pkl`, main.py:54-63) name its top-level
    packages (`helpers.BaseReader.BaseReader`); they hold the same attributes, so they load as this package's reader."""

    def find_class(self, module, name):
        if module.split('.')[0] in ('helpers', 'models', 'utils') and not module.startswith(__package__):
            try:
                return getattr(importlib.import_module('.' + module, package=__package__), name)
            except (ImportError, AttributeError):
                pass
        return super().find_class(module, name)


def load_corpus(path):
    with open(path, 'rb') as f:
        return _CorpusUnpickler(f).load()


def save_corpus(corpus, path):
    """Write the corpus cache so that BOTH programs can load it (reference main.py:54-63): the reader object is pickled
    under the reference's class path (`helpers.<Reader>.<Reader>`, plain attributes: `n_users`, `n_items`, `data_df`,
    `train_clicked_set`, `residual_clicked_set`, ...), which `load_corpus` maps back to this package's class.  Written to a
    temporary file and renamed, so that a concurrent reader never sees half a pickle."""
    import types
    cls = type(corpus)
    mod_name = 'helpers.' + cls.__name__
    stand_in = type(cls.__name__, (object,), {})
    stand_in.__module__ = mod_name
    alias_pkg, alias_mod = types.ModuleType('helpers'), types.ModuleType(mod_name)
    setattr(alias_mod, cls.__name__, stand_in)
    setattr(alias_pkg, cls.__name__, alias_mod)
    saved = {k: sys.modules.get(k) for k in ('helpers', mod_name)}
    obj = stand_in.__new__(stand_in)
    obj.__dict__.update(corpus.__dict__)
    tmp = '{}.tmp.{}'.format(path, os.getpid())
    try:
        sys.modules['helpers'], sys.modules[mod_name] = alias_pkg, alias_mod
        with open(tmp, 'wb') as f:
            pickle.dump(obj, f)
        os.replace(tmp, path)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if os.path.exists(tmp):
            os.remove(tmp)


def parse_global_args(parser):
    parser.add_argument('--gpu', type=str, default='0', help='Set CUDA_VISIBLE_DEVICES, default for CPU only')
    parser.add_argument('--verbose', type=int, default=logging.INFO, help='Logging Level, 0, 10, ..., 50')
    parser.add_argument('--log_file', type=str, default='', help='Logging file path')
    parser.add_argument('--random_seed', type=int, default=3407,
                        help='Random seed of numpy and pytorch, and 3407 is all you need')
    parser.add_argument('--load', type=int, default=0, help='Whether load model and continue to train')
    parser.add_argument('--train', type=int, default=1, help='To train the model or not.')
    parser.add_argument('--regenerate', type=int, default=1, help='Whether to regenerate intermediate files')
    parser.add_argument('--eval_precision', type=int, default=2,
                        help='Full-ranking eval: 0 fp32 FMA tiles; 1 bf16 tensor cores (looser); 2 split-bf16 tensor cores '
                             'with an exact re-check -- the ranks of 0 at tensor-core speed (embedding 64 / 128, else 0)')
    return parser


def build_args(argv=None):
    """Two-stage parse of main.py:104-122.  Returns (init_args, args, classes)."""
    init_parser = argparse.ArgumentParser(description='Model')
    init_parser.add_argument('--model_name', type=str, default='BPRMF', help='Choose a model to run.')
    init_parser.add_argument('--reader_name', type=str, default=None, help='Choose a reader object.')
    init_parser.add_argument('--runner_name', type=str, default=None, help='Choose a runner object.')
    init_args, _ = init_parser.parse_known_args(argv)
    model_class = resolve('model', init_args.model_name)
    reader_name = model_class.reader if init_args.reader_name is None else init_args.reader_name
    runner_name = model_class.runner if init_args.runner_name is None else init_args.runner_name
    reader_class, runner_class = resolve('helper', reader_name), resolve('helper', runner_name)

    parser = argparse.ArgumentParser(description='')
    parser = parse_global_args(parser)
    parser = reader_class.parse_reader_args(parser)
    parser = runner_class.parse_runner_args(parser)
    parser = model_class.parse_model_args(parser)
    args, _ = parser.parse_known_args(argv)

    log_args = [init_args.model_name, args.dataset, str(args.random_seed)]
    for arg in ['lr', 'l2'] + model_class.extra_log_args:
        log_args.append(arg + '=' + str(getattr(args, arg)))
    log_file_name = '__'.join(log_args).replace(' ', '__')
    if args.log_file == '':
        args.log_file = '../log/{}/{}.txt'.format(init_args.model_name, log_file_name)
    if args.model_path == '':
        args.model_path = '../model/{}/{}.pt'.format(init_args.model_name, log_file_name)
    return init_args, args, (model_class, reader_class, runner_class, reader_name)


def default_args(model_class, **overrides):
    """The reference's defaults for a model (global + reader + runner + model flags), for programmatic use."""
    from .helpers.BaseReader import BaseReader
    from .helpers.BaseRunner import BaseRunner
    parser = argparse.ArgumentParser()
    parser = parse_global_args(parser)
    parser = BaseReader.parse_reader_args(parser)
    parser = BaseRunner.parse_runner_args(parser)
    parser = model_class.parse_model_args(parser)
    args, _ = parser.parse_known_args([])
    args.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    for key, value in overrides.items():
        setattr(args, key, value)
    return args


def run(args, model_class, reader_class, runner_class, reader_name):
    """main.py:36-85."""
    logging.info('-' * 45 + ' BEGIN: ' + utils.get_time() + ' ' + '-' * 45)
    exclude = ['check_epoch', 'log_file', 'model_path', 'path', 'pin_memory', 'load',
               'regenerate', 'sep', 'train', 'verbose', 'metric', 'test_epoch', 'buffer']
    logging.info(utils.format_arg_str(args, exclude_lst=exclude))

    utils.init_seed(args.random_seed)

    world = int(os.environ.get('WORLD_SIZE', 1))
    if world == 1:
        os.environ['CUDA_VISIBLE_DEVICES'] = args.gpu
    if args.gpu == '' or not torch.cuda.is_available():
        raise RuntimeError('whisprrec_b200 needs a CUDA (sm_100) device: it has no CPU path. '
                           'Run the reference for a CPU run.')
    peers = None
    if world > 1:
        # torchrun --nproc-per-node G main.py ...: one rank per GPU of the box, tables row-sharded over them
        import torch.distributed as dist
        from . import sharded
        args.device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
        torch.cuda.set_device(args.device)
        dist.init_process_group('nccl', device_id=args.device)
        peers = sharded.PeerGroup(args.device)
        if peers.rank != 0:
            logging.getLogger().setLevel(logging.WARNING)
    else:
        args.device = torch.device('cuda')
    logging.info('Device: {}'.format(args.device))

    corpus_path = os.path.join(args.path, args.dataset, reader_name + '.pkl')
    if not args.regenerate and os.path.exists(corpus_path):
        logging.info('Load corpus from {}'.format(corpus_path))
        corpus = load_corpus(corpus_path)
    else:
        corpus = reader_class(args)
        if peers is None or peers.rank == 0:     # one writer (every rank builds the same corpus from the same file)
            logging.info('Save corpus to {}'.format(corpus_path))
            try:
                save_corpus(corpus, corpus_path)
            except OSError as e:                 # read-only data dir: the cache is an optimisation only
                logging.info('corpus not cached: {}'.format(e))
        if peers is not None:
            peers.host_sync()

    model = model_class(args, corpus).to(args.device)
    logging.info('#params: {}'.format(model.count_variables()))
    logging.info(model)

    data_dict = {phase: model_class.Dataset(model, corpus, phase) for phase in ('train', 'dev', 'test')}
    runner = runner_class(args)
    if peers is not None:
        model.fuse()
        model.shard(peers)                       # every rank holds the same seed-initialised tables at this point
    if args.load > 0:
        model.load_model()
    if args.train > 0:
        runner.train(data_dict)
    eval_res = runner.print_res(data_dict['test'])
    logging.info(os.linesep + 'Test After Training: ' + eval_res)
    model.actions_after_train()
    logging.info(os.linesep + '-' * 45 + ' END: ' + utils.get_time() + ' ' + '-' * 45)
    if peers is not None:
        import torch.distributed as dist
        model.unshard()                          # leave the full tables behind for whoever holds the model object
        peers.host_sync()
        dist.destroy_process_group()
    return model, runner, data_dict


def main(argv=None):
    init_args, args, (model_class, reader_class, runner_class, reader_name) = build_args(argv)
    utils.check_dir(args.log_file)
    logging.basicConfig(filename=args.log_file, level=args.verbose)
    logging.getLogger().addHandler(logging.StreamHandler(sys.stdout))
    logging.info(init_args)
    return run(args, model_class, reader_class, runner_class, reader_name)


if __name__ == '__main__':
    main()

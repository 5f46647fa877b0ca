"""Command line with the reference's flag grammar and messages (/root/reference/src/tezip.py:10-101).

  python -m tezip_b200.tezip -c MODEL_DIR IMG_DIR OUT_DIR -p P (-w W | -t T) -m {abs,rel,absrel,pwrel} -b V [V2] [-v] [-n]
  python -m tezip_b200.tezip -u MODEL_DIR COMP_DIR OUT_DIR [-v]

-l (training) is outside the scope of this build; -f (force CPU) is rejected: there is no CPU path.
The flags and the validation order are data (FLAGS, COMPRESS_RULES) walked by small helpers, so that the grammar can
be compared with the reference's at a glance.
"""
import argparse
import sys

BOUND_MODES = {'abs': 1, 'rel': 1, 'pwrel': 1, 'absrel': 2}         # mode -> number of -b values (tezip.py:66-70)

# (short, long, argparse keywords) in the reference's order (tezip.py:88-100)
FLAGS = (
    ('-l', '--learn', dict(type=str, nargs=2, metavar=('model', 'dir'))),
    ('-c', '--compress', dict(type=str, nargs=3, metavar=('model', 'dir', 'file'))),
    ('-u', '--uncompress', dict(type=str, nargs=3, metavar=('model', 'file', 'dir'))),
    ('-p', '--preprocess', dict(type=int, nargs=1, metavar='warm_up_num')),
    ('-w', '--window', dict(type=int, nargs=1, metavar='window_size')),
    ('-t', '--threshold', dict(type=float, nargs=1, metavar='MSE_threshold')),
    ('-m', '--mode', dict(type=str, nargs=1, metavar='mode')),
    ('-b', '--bound', dict(type=float, nargs='*', metavar='value', default=None)),
    ('-f', '--force', dict(action='store_true')),
    ('-v', '--verbose', dict(action='store_true')),
    ('-n', '--no_entropy', dict(action='store_false')),
)

SWP_DWP_HINT = 'Select window size for SWP and MSE threshold for DWP.'
HELP_HINT = 'Command to check the options is -h or --help'

# (predicate over the parsed args that means "invalid", message lines) checked in the reference's order (tezip.py:39-70)
COMPRESS_RULES = (
    (lambda a: a.preprocess is None,
     ('Please specify the -p or --preprocess option!', 'warm up num.')),
    (lambda a: a.window is None and a.threshold is None,
     ('Please specify the window size(-w or --window) or MSE threshold(-t or --threshold) option!', SWP_DWP_HINT)),
    (lambda a: a.window is not None and a.threshold is not None,
     ('Please select only one of window size(-w or --window) or MSE threshold(-t or --threshold)!', SWP_DWP_HINT)),
    (lambda a: a.mode is None or a.mode[0] not in BOUND_MODES,
     ('Please specify the -m or --mode correctly!', "'abs' or 'rel' or 'absrel' or 'pwrel'.")),
    (lambda a: not a.bound,
     ('Please specify the -b or --bound option!', 'error bound value.')),
    (lambda a: len(a.bound) != BOUND_MODES[a.mode[0]],
     ("If the -m or --mode is 'abs' or 'rel' or 'pwrel', enter one for -b or --bound. : value",
      "If the -m or --mode is 'absrel', enter two in -b or --bound. : abs_value rel_value")),
)


def gpu_available():
    import torch
    from . import _lib
    return torch.cuda.is_available() and _lib.load().tz_device_count() > 0


def _fail(*lines):
    print('ERROR')
    for line in lines:
        print(line)
    return 2


def _first(opt):
    return None if opt is None else opt[0]


def _compress(arg, on_gpu):
    print('compress mode')
    for k, (invalid, lines) in enumerate(COMPRESS_RULES):
        if k == 4:
            print(arg.mode[0])                                      # the reference echoes the mode once it is valid
        if invalid(arg):
            return _fail(*lines)
    from . import compress
    model_dir, image_dir, out_dir = arg.compress
    compress.run(model_dir, image_dir, out_dir, arg.preprocess[0], _first(arg.window), _first(arg.threshold),
                 arg.mode[0], arg.bound, on_gpu, arg.verbose, arg.no_entropy)
    return 0


def _uncompress(arg, on_gpu):
    print('uncompress mode')
    from . import decompress
    model_dir, comp_dir, out_dir = arg.uncompress
    decompress.run(model_dir, comp_dir, out_dir, on_gpu, arg.verbose)
    return 0


def _learn(arg, on_gpu):
    print('train mode')
    return _fail('training is not part of tezip_b200; train with the reference and convert the weights.')


def main(arg):
    if arg.force:
        return _fail('-f/--force (CPU mode) is not available: tezip_b200 runs on B200 GPUs only.')
    on_gpu = gpu_available()
    print('GPU MODE' if on_gpu else 'CPU MODE')
    actions = [(arg.learn, _learn), (arg.compress, _compress), (arg.uncompress, _uncompress)]
    chosen = [fn for val, fn in actions if val is not None]
    if len(chosen) > 1:
        return _fail('Please select only one of learn or compress or uncompress.', HELP_HINT)
    if not chosen:
        return _fail('Please mode select!', 'learn or compress or uncompress.', HELP_HINT)
    return chosen[0](arg, on_gpu)


def build_parser():
    parser = argparse.ArgumentParser(prog='TEZIP', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for short, long_, kw in FLAGS:
        parser.add_argument(short, long_, dest=long_[2:], **kw)
    return parser


if __name__ == '__main__':
    sys.exit(main(build_parser().parse_args()))

"""Command line of the reference (/root/reference/src/tezip.py:10-101), flag for flag.

  python -m tezip_b200.tezip -c MODEL_DIR IMG_DIR OUT_DIR -p P (-w W | -t T) -m {abs,rel,absrel,pwrel} -b V [V2] [-v] [-n]
  python -m tezip_b200.tezip -u MODEL_DIR COMP_DIR OUT_DIR [-v]

-l (training) is outside the scope of this build; -f (force CPU) is rejected: there is no CPU path.
"""
import argparse
import sys


def gpu_available():
    import torch
    from . import _lib
    return torch.cuda.is_available() and _lib.load().tz_device_count() > 0


def main(arg):
    if arg.force:                                                   # tezip.py:12-13
        print('ERROR')
        print('-f/--force (CPU mode) is not available: tezip_b200 runs on B200 GPUs only.')
        return 2
    GPU_flag = gpu_available()                                      # tezip.py:16-21
    print('GPU MODE' if GPU_flag else 'CPU MODE')
    n_sel = sum(x is not None for x in (arg.learn, arg.compress, arg.uncompress))
    if n_sel > 1:                                                   # tezip.py:28-31
        print('ERROR')
        print('Please select only one of learn or compress or uncompress.')
        print('Command to check the options is -h or --help')
        return 2
    if arg.learn is not None:                                       # tezip.py:33-35
        print('train mode')
        print('ERROR')
        print('training is not part of tezip_b200; train with the reference and convert the weights.')
        return 2
    if arg.compress is not None:                                    # tezip.py:37-74
        print('compress mode')
        if arg.preprocess is None:
            print('ERROR')
            print('Please specify the -p or --preprocess option!')
            print('warm up num.')
            return 2
        if arg.window is None and arg.threshold is None:
            print('ERROR')
            print('Please specify the window size(-w or --window) or MSE threshold(-t or --threshold) option!')
            print('Select window size for SWP and MSE threshold for DWP.')
            return 2
        if arg.window is not None and arg.threshold is not None:
            print('ERROR')
            print('Please select only one of window size(-w or --window) or MSE threshold(-t or --threshold)!')
            print('Select window size for SWP and MSE threshold for DWP.')
            return 2
        if arg.mode is None or arg.mode[0] not in ('abs', 'rel', 'absrel', 'pwrel'):
            print('ERROR')
            print('Please specify the -m or --mode correctly!')
            print('\'abs\' or \'rel\' or \'absrel\' or \'pwrel\'.')
            return 2
        print(arg.mode[0])
        if arg.bound is None or len(arg.bound) == 0:
            print('ERROR')
            print('Please specify the -b or --bound option!')
            print('error bound value.')
            return 2
        if not ((arg.mode[0] in ('abs', 'rel', 'pwrel') and len(arg.bound) == 1) or
                (arg.mode[0] == 'absrel' and len(arg.bound) == 2)):
            print('ERROR')
            print('If the -m or --mode is \'abs\' or \'rel\' or \'pwrel\', enter one for -b or --bound. : value')
            print('If the -m or --mode is \'absrel\', enter two in -b or --bound. : abs_value rel_value')
            return 2
        from . import compress
        compress.run(arg.compress[0], arg.compress[1], arg.compress[2], arg.preprocess[0],
                     arg.window[0] if arg.window is not None else None,
                     arg.threshold[0] if arg.threshold is not None else None,
                     arg.mode[0], arg.bound, GPU_flag, arg.verbose, arg.no_entropy)
        return 0
    if arg.uncompress is not None:                                  # tezip.py:76-78
        print('uncompress mode')
        from . import decompress
        decompress.run(arg.uncompress[0], arg.uncompress[1], arg.uncompress[2], GPU_flag, arg.verbose)
        return 0
    print('ERROR')                                                  # tezip.py:80-84
    print('Please mode select!')
    print('learn or compress or uncompress.')
    print('Command to check the options is -h or --help')
    return 2


def build_parser():
    """tezip.py:88-100, verbatim flag grammar."""
    parser = argparse.ArgumentParser(prog='TEZIP', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-l', '--learn', type=str, nargs=2, metavar=('model', 'dir'), dest='learn')
    parser.add_argument('-c', '--compress', type=str, nargs=3, metavar=('model', 'dir', 'file'), dest='compress')
    parser.add_argument('-u', '--uncompress', type=str, nargs=3, metavar=('model', 'file', 'dir'), dest='uncompress')
    parser.add_argument('-p', '--preprocess', type=int, nargs=1, metavar=('warm_up_num'), dest='preprocess')
    parser.add_argument('-w', '--window', type=int, nargs=1, metavar=('window_size'), dest='window')
    parser.add_argument('-t', '--threshold', type=float, nargs=1, metavar=('MSE_threshold'), dest='threshold')
    parser.add_argument('-m', '--mode', type=str, nargs=1, metavar=('mode'), dest='mode')
    parser.add_argument('-b', '--bound', type=float, nargs='*', metavar=('value'), dest='bound', default=None)
    parser.add_argument('-f', '--force', action='store_true')
    parser.add_argument('-v', '--verbose', action='store_true')
    parser.add_argument('-n', '--no_entropy', action='store_false')
    return parser


if __name__ == '__main__':
    sys.exit(main(build_parser().parse_args()))

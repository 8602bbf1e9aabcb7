"""Host-side argument handling for the grouping stage -- behaviour of videotofaces.prep
(/root/reference/src/videotofaces/prep.py:18-120) for the flags the hot path reads:
--clusters, --ref-dir, -o/--out-dir, -m/--mode, --group-mode.  Messages match the reference.
"""
import os
import os.path as osp


def _say(*a):
    """Console output comes from one process: under torchrun (one process per GPU) every rank runs the same argument
    checks, rank 0 alone reports them."""
    if int(os.environ.get('RANK', '0')) == 0:
        print(*a)


IMG_EXTENSIONS = ('.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.pgm', '.tif', '.tiff', '.webp')

_CHOICES = {
    'mode': ['full', 'detection', 'grouping'],
    'style': ['live', 'anime'],
    'group_mode': ['clustering', 'classification'],
    'video_reader': ['opencv', 'decord'],
}
_MODELS = {
    'live': dict(det_model=['default', 'yolo', 'mtcnn'], enc_model=['default', 'facenet_vgg', 'facenet_casia']),
    'anime': dict(det_model=['default', 'rcnn'], enc_model=['default', 'vit_b', 'vit_l']),
}


def _is_image(entry):
    return entry.is_file() and entry.name.lower().endswith(IMG_EXTENSIONS)


def get_img_paths(target_dir):
    return sorted(e.path for e in os.scandir(target_dir) if _is_image(e))


def _known(val, arg_name, options):
    if val in options:
        return True
    _say('ERROR: unknown %s. Available options are %s' % (arg_name, ', '.join('"' + v + '"' for v in options)))
    return False


def validate_args(mode, input_path, out_dir, style, group_mode, video_reader, det_model, enc_model):
    """prep.py:18-45."""
    if not _known(mode, 'mode', _CHOICES['mode']):
        return False
    problems = []
    if input_path and not osp.exists(input_path):
        problems.append('ERROR: specified input_path doesn\'t exist. Please provide a valid path to a file, '
                        'a directory with files, or a .txt file with full paths inside')
    if out_dir and not osp.isdir(out_dir):
        problems.append('ERROR: specified out_dir doesn\'t exist or isn\'t a directory. Please provide a valid path to a directory')
    if not input_path and mode != 'grouping':
        problems.append('ERROR: please specify input_path')
    if not input_path and mode == 'grouping' and not out_dir:
        problems.append('ERROR: for grouping, please specify either out_dir or the same input_path used during detection')
    for p in problems:
        _say(p)
    ok = not problems
    for name, val in (('style', style), ('group_mode', group_mode), ('video_reader', video_reader)):
        ok = ok and _known(val, name, _CHOICES[name])
    if style in _MODELS:
        ok = ok and _known(det_model, 'det_model', _MODELS[style]['det_model'])
        ok = ok and _known(enc_model, 'enc_model', _MODELS[style]['enc_model'])
    return ok


def get_clusters(c):
    """prep.py:48-66: None -> 2..8, int -> [c], 'a,b,c' -> sorted set, 'A-B' -> range."""
    if not c:
        return list(range(2, 9))
    if isinstance(c, int) and c > 0:
        return [c]
    if isinstance(c, str):
        if ',' in c:
            parts = c.split(',')
            if all(p.isdigit() for p in parts):
                return sorted({int(p) for p in parts})
        parts = c.split('-')
        if len(parts) == 2 and parts[0].isdigit() and parts[1].isdigit():
            lo, hi = int(parts[0]), int(parts[1])
            if 0 < lo < hi:
                return list(range(lo, hi + 1))
    _say('ERROR: incorrent value for clusters. Please specify a natural number or a string either as an '
          'enumeration "C1,C2,C3,C4" or a range "A-B" where 0 < A < B')
    return None


def get_class_ref(ref_dir, out_dir):
    """prep.py:69-105: [(class_name, [image paths])] sorted by class, classes without images dropped."""
    explanation = ('Please prepare a directory with 1 or more subfolders representing groups, each with 1 or more '
                   'reference images inside')
    if not ref_dir:
        guess = osp.join(out_dir, 'ref')
        if not osp.isdir(guess):
            _say('ERROR: for group_mode="classification", ref_dir needs to be specified')
            _say(explanation)
            return None
        _say('NOTE: ref_dir is unspecified, but found "ref" folder inside out_dir. Will search for reference images there')
        ref_dir = guess
    if not osp.isdir(ref_dir):
        _say('ERROR: specified ref_dir doesn\'t exist or isn\'t a directory. Please provide a valid path to a directory')
        return None
    classes = sorted(e.name for e in os.scandir(ref_dir) if e.is_dir())
    if not classes:
        _say('ERROR: specified ref_dir doesn\'t contain any subfolders')
        _say(explanation)
        return None
    refs, warnings = [], []
    for c in classes:
        imgs = get_img_paths(osp.join(ref_dir, c))
        if imgs:
            refs.append((c, imgs))
        else:
            warnings.append('WARNING: ref_dir\'s subfolder "%s" doesn\'t contain any images. During classification, '
                            'this class will be ignored' % c)
    if not refs:
        _say('ERROR: none of the ref_dir\'s subfolders contain any images')
        _say('Supported extensions are: %s' % ', '.join(IMG_EXTENSIONS))
        return None
    for w in warnings:
        _say(w)
    return refs


def get_paths_for_grouping(out_dir):
    """prep.py:108-120: <out_dir>/faces first, then out_dir itself."""
    for tdir in (osp.join(out_dir, 'faces'), out_dir):
        paths = get_img_paths(tdir) if osp.isdir(tdir) else []
        if paths:
            _say('Found %u images at: %s' % (len(paths), tdir))
            return paths
    _say('ERROR: no image files for grouping found at: %s' % out_dir)
    return None

"""Golden fixtures of the step before compositing, from the LIVE
reference (vkit-x/vkit @ 98ada2d under /root/reference, cv2 4.13.0.92, numpy 2.3.5).  Run in the
build container only:

    PYTHONPATH=/root/repo python tests/golden/make_golden_f4.py

  combiner      ImageCombinerEngine.run (engine/image/combiner.py:91-345) on synthetic textures
                written as PNG files into a temporary metas folder; the rectangles the engine
                pastes are logged by wrapping fill_np_edge_mask, the rng state after the run is
                recorded.
  glyph         build_char_glyph (engine/font/freetype.py:136-221) on synthetic coverage bitmaps
                with a stand-in glyph slot: trimming, paddings, ascent, gamma-corrected alpha.
  text_line     render_char_glyphs_in_text_line (freetype.py:314-380), default and LCD branches.
`freetype` itself is absent here (no FT_Face, no font files): the module is imported with an inert
stand-in, only its NumPy / cv2 code runs.  Inputs are regenerated from seeds by tests/common.py.
"""
import json
import os
import pathlib
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

import make_golden as mg  # noqa: E402  (loads the reference through oracle/refshim)


class _InertModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return type(name, (), {})


sys.modules.setdefault('freetype', _InertModule('freetype'))

import cv2  # noqa: E402
from vkit.element import Box, Image, ScoreMap  # noqa: E402
import vkit.element.image as ref_image_module  # noqa: E402
import vkit.engine.image.combiner as ref_combiner  # noqa: E402
from vkit.engine.image.type import ImageEngineRunConfig  # noqa: E402
import vkit.engine.font.freetype as ref_freetype  # noqa: E402
from vkit.engine.font.type import (CharBox, FontEngineRunConfigGlyphSequence,  # noqa: E402
                                   FontEngineRunConfigStyle, FontGlyphInfo)

from common import f4_glyph_bitmaps, f4_textures  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES, ARRAYS = [], {}


class _PathIo:
    """iolite as the two modules use it: paths in, pathlib paths out."""

    @staticmethod
    def file(path, exists=False, expandvars=False):
        return pathlib.Path(path)

    @staticmethod
    def folder(path, exists=False, expandvars=False, touch=False):
        return pathlib.Path(path)


def combiner_cases():
    ref_image_module.io = _PathIo
    ref_combiner.io = _PathIo
    ref_combiner.read_json_file = lambda path: json.load(open(path))
    specs = [
        # (textures seed, count, size range, canvas, config overrides, rng seeds of successive runs)
        (7001, 6, (40, 90), (128, 160), {}, [1, 2, 3]),
        (7002, 9, (30, 70), (200, 333), {'prob_use_only_the_anchor_image': 0.0}, [4, 5]),
        (7003, 5, (60, 200), (97, 61), {'prob_rotate_image': 1.0}, [6]),
        (7004, 8, (25, 60), (256, 256), {'prob_use_only_the_anchor_image': 0.0,
                                         'enable_cache': True}, [7, 8, 9]),
        (7005, 12, (100, 300), (1024, 1024), {'prob_use_only_the_anchor_image': 0.3}, [10, 11]),
        (7006, 4, (20, 40), (64, 300), {'gaussian_blur_kernel_size': 7,
                                        'init_segment_width_min_ratio': 0.1}, [12]),
        (7007, 3, (300, 400), (90, 120), {}, [13]),
    ]
    for k, (tex_seed, count, (smin, smax), (height, width), overrides, seeds) in enumerate(specs):
        textures = f4_textures(tex_seed, count, smin, smax)
        with tempfile.TemporaryDirectory() as folder:
            os.makedirs(os.path.join(folder, 'image'))
            metas = []
            for name, mat, mean, std in textures:
                cv2.imwrite(os.path.join(folder, 'image', name), mat[:, :, ::-1])
                metas.append({'image_file': name, 'grayscale_mean': mean, 'grayscale_std': std})
            with open(os.path.join(folder, 'metas.json'), 'w') as fout:
                json.dump(metas, fout)
            init_config = ref_combiner.ImageCombinerEngineInitConfig(image_meta_folder=folder)
            for key, value in overrides.items():
                if key == 'gaussian_blur_kernel_size':
                    # a class attribute without annotation in the reference (combiner.py:81)
                    ref_combiner.ImageCombinerEngineInitConfig.gaussian_blur_kernel_size = value
                else:
                    setattr(init_config, key, value)
            engine = ref_combiner.ImageCombinerEngine(init_config)
            logged = []
            original = ref_combiner.ImageCombinerEngine.fill_np_edge_mask.__func__

            def logging_fill(cls, np_edge_mask, height, width, gaussian_blur_half_kernel_size, up,
                             down, left, right):
                logged.append([up, down, left, right])
                return original(cls, np_edge_mask, height, width, gaussian_blur_half_kernel_size,
                                up, down, left, right)

            ref_combiner.ImageCombinerEngine.fill_np_edge_mask = classmethod(logging_fill)
            try:
                runs = []
                for seed in seeds:
                    del logged[:]
                    rng = np.random.default_rng(seed)
                    image = engine.run(ImageEngineRunConfig(height=height, width=width), rng)
                    run = {'rng_seed': seed, 'rects': [list(r) for r in logged],
                           'sha': mg.sha(image.mat), 'rng_after': int(rng.integers(0, 2**31))}
                    runs.append(run)
                    if height * width <= 70000 and seed == seeds[0]:
                        ARRAYS[f'cb{k:02d}/image'] = image.mat
            finally:
                ref_combiner.ImageCombinerEngine.fill_np_edge_mask = classmethod(original)
                ref_combiner.ImageCombinerEngineInitConfig.gaussian_blur_kernel_size = 5
        CASES.append({'id': f'cb{k:02d}', 'kind': 'combiner', 'textures_seed': tex_seed,
                      'textures_count': count, 'size_range': [smin, smax],
                      'canvas': [height, width], 'config': overrides, 'runs': runs})


class _Slot:
    def __init__(self, top, left, advance_x):
        self.bitmap_top, self.bitmap_left = top, left
        self.advance = types.SimpleNamespace(x=advance_x, y=0)


def _glyph_config(gamma):
    info = FontGlyphInfo(tags=['t'], ascent_plus_pad_up_min_to_font_size_ratio=0.8,
                         height_min_to_font_size_ratio=1.0, width_min_to_font_size_ratio=0.6)
    variant = types.SimpleNamespace(
        char_to_tags={'a': ['t']},
        font_glyph_info_collection=types.SimpleNamespace(tag_to_font_glyph_info={'t': info}))
    return types.SimpleNamespace(
        style=FontEngineRunConfigStyle(glyph_color_gamma=gamma, glyph_color=(30, 60, 90)),
        font_variant=variant, glyph_sequence=FontEngineRunConfigGlyphSequence.HORI_DEFAULT,
        height=32, width=400)


def glyph_and_text_line_cases():
    specs = [(8001, 14, False, 1.0), (8002, 10, False, 1.7), (8003, 12, True, 1.0),
             (8004, 9, True, 0.6), (8005, 16, False, 0.45)]
    for k, (seed, count, lcd, gamma) in enumerate(specs):
        config = _glyph_config(gamma)
        char_glyphs, metrics = [], []
        for j, (bitmap, top, left, advance_x) in enumerate(f4_glyph_bitmaps(seed, count, lcd)):
            glyph = ref_freetype.build_char_glyph(config, 'a', _Slot(top, left, advance_x), bitmap)
            char_glyphs.append(glyph)
            metrics.append({'height': glyph.height, 'width': glyph.width, 'ascent': glyph.ascent,
                            'pad_up': glyph.pad_up, 'pad_down': glyph.pad_down,
                            'pad_left': glyph.pad_left, 'pad_right': glyph.pad_right,
                            'sha_image': mg.sha(glyph.image.mat),
                            'sha_alpha': mg.sha(glyph.score_map.mat) if glyph.score_map else None})
        # place the glyphs on a line, neighbours overlapping by two columns, varying baselines
        rng = np.random.default_rng(seed + 1)
        line_height = max(g.height for g in char_glyphs) + 6
        boxes, x = [], 1
        for glyph in char_glyphs:
            up = int(rng.integers(0, line_height - glyph.height + 1))
            boxes.append([up, up + glyph.height - 1, x, x + glyph.width - 1])
            x += max(1, glyph.width - 2)
        line_width = x + 30
        char_boxes = [CharBox(char='a', box=Box(up=b[0], down=b[1], left=b[2], right=b[3]))
                      for b in boxes]
        image, mask, score_map, _ = ref_freetype.render_char_glyphs_in_text_line(
            config.style, line_height, line_width, char_glyphs, char_boxes)
        cid = f'tl{k:02d}'
        CASES.append({'id': cid, 'kind': 'text_line', 'seed': seed, 'count': count, 'lcd': lcd,
                      'gamma': gamma, 'glyph_color': [30, 60, 90], 'metrics': metrics,
                      'boxes': boxes, 'line_shape': [line_height, line_width],
                      'sha': {'image': mg.sha(image.mat), 'mask': mg.sha(mask.mat),
                              'score_map': mg.sha(score_map.mat) if score_map else None}})
        ARRAYS[f'{cid}/image'] = image.mat
        if score_map:
            ARRAYS[f'{cid}/score_map'] = score_map.mat


def main():
    combiner_cases()
    glyph_and_text_line_cases()
    with open(os.path.join(HERE, 'f4_cases.json'), 'w') as fout:
        json.dump({'reference': 'vkit-x/vkit@98ada2d', 'cv2': cv2.__version__,
                   'numpy': np.__version__, 'cases': CASES}, fout, indent=1)
    np.savez_compressed(os.path.join(HERE, 'f4_arrays.npz'), **ARRAYS)
    print(len(CASES), 'cases;', sum(v.nbytes for v in ARRAYS.values()) / 1e6, 'MB raw arrays')


if __name__ == '__main__':
    main()

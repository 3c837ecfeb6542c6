"""The step before compositing: background synthesis and the glyph
atlas.  CPU part: the oracle and the host logic against fixtures of the live reference
(tests/golden/make_golden_f4.py).  GPU part: the device path against the same fixtures."""
import numpy as np
import pytest

from common import f4_array, f4_cases, f4_glyph_bitmaps, f4_textures, sha

COMBINER = f4_cases('combiner')
TEXT_LINES = f4_cases('text_line')


def _combiner_for(case):
    from vkit_b200.background import ImageCombiner, ImageCombinerConfig, Texture
    from vkit_b200.element import Image
    textures = f4_textures(case['textures_seed'], case['textures_count'], *case['size_range'])
    config = ImageCombinerConfig(**case['config'])
    combiner = ImageCombiner([Texture(name, Image(mat=mat), mean, std)
                              for name, mat, mean, std in textures], config)
    return combiner, textures


@pytest.mark.parametrize('case', COMBINER, ids=[c['id'] for c in COMBINER])
def test_combiner_walk_matches_reference(case):
    """Same rectangles in the same order and the same generator state afterwards as
    ImageCombinerEngine.run -- including enable_cache state carried across runs."""
    combiner, _ = _combiner_for(case)
    height, width = case['canvas']
    for run in case['runs']:
        rng = np.random.default_rng(run['rng_seed'])
        placements = combiner.plan(height, width, combiner.sample_candidates(rng), rng)
        assert [[p.up, p.down, p.left, p.right] for p in placements] == run['rects']
        assert int(rng.integers(0, 2**31)) == run['rng_after']


def _oracle_pastes(combiner, textures, placements):
    """Texture arrays in the orientation the walk chose, rotated by the oracle's own warp."""
    from oracle import vkit_port as port
    by_name = {name: mat for name, mat, _, _ in textures}
    pastes = []
    for p in placements:
        mat = by_name[combiner.textures[p.texture].name]
        if p.rotated:
            trans_mat, dsize = port.affine_state('rotate', {'angle': 90}, mat.shape[:2])
            mat = port.affine_apply(mat, trans_mat, dsize)
        pastes.append((mat, p.up, p.down, p.left, p.right))
    return pastes


@pytest.mark.parametrize('case', [c for c in COMBINER if c['canvas'][0] * c['canvas'][1] < 100000],
                         ids=lambda c: c['id'])
def test_oracle_combiner_matches_reference(case):
    from oracle import vkit_port as port
    combiner, textures = _combiner_for(case)
    height, width = case['canvas']
    ksize = case['config'].get('gaussian_blur_kernel_size', 5)
    for use_cv2 in (False, True):
        port.use_cv2(use_cv2)
        try:
            combiner._cached_orientation.clear()
            for run in case['runs']:
                rng = np.random.default_rng(run['rng_seed'])
                placements = combiner.plan(height, width, combiner.sample_candidates(rng), rng)
                got = port.combiner_compose((height, width),
                                            _oracle_pastes(combiner, textures, placements), ksize)
                assert sha(got) == run['sha'], (case['id'], run['rng_seed'], use_cv2)
        finally:
            port.use_cv2(False)


@pytest.mark.parametrize('case', TEXT_LINES, ids=[c['id'] for c in TEXT_LINES])
def test_oracle_glyphs_and_text_line_match_reference(case):
    from oracle import vkit_port as port
    glyphs = []
    for (bitmap, _, _, _), metrics in zip(f4_glyph_bitmaps(case['seed'], case['count'], case['lcd']),
                                         case['metrics']):
        planes = port.glyph_planes(bitmap, case['gamma'])
        assert sha(planes[0]) == metrics['sha_image']
        if not case['lcd']:
            assert sha(planes[2]) == metrics['sha_alpha']
        glyphs.append(planes)
    image, mask, score_map = port.render_text_line(case['glyph_color'], case['line_shape'], glyphs,
                                                   case['boxes'])
    assert sha(image) == case['sha']['image'] and sha(mask) == case['sha']['mask']
    if case['lcd']:
        assert score_map is None
    else:
        assert sha(score_map) == case['sha']['score_map']


@pytest.mark.parametrize('case', TEXT_LINES, ids=[c['id'] for c in TEXT_LINES])
def test_atlas_metrics_and_tables_match_reference(case):
    """GlyphAtlas.add derives build_char_glyph's paddings / ascent; the 256-entry gamma tables
    reproduce the reference's per-pixel np.power on every glyph (host logic, no device)."""
    from vkit_b200.compositing import GlyphAtlas, trim_glyph_bitmap
    atlas = GlyphAtlas()
    alpha_lut, lcd_lut = GlyphAtlas.gamma_tables(case['gamma'])
    bitmaps = f4_glyph_bitmaps(case['seed'], case['count'], case['lcd'])
    for j, ((bitmap, top, left, advance_x), metrics) in enumerate(zip(bitmaps, case['metrics'])):
        glyph = atlas.add(j, bitmap, gamma=case['gamma'], bitmap_top=top, bitmap_left=left,
                          advance_x=advance_x)
        for key in ('height', 'width', 'ascent', 'pad_up', 'pad_down', 'pad_left', 'pad_right'):
            assert getattr(glyph, key) == metrics[key], (j, key)
        trimmed = trim_glyph_bitmap(bitmap)[0]
        assert sha(trimmed) == metrics['sha_image']
        if not case['lcd']:
            assert sha(alpha_lut[trimmed]) == metrics['sha_alpha']


# ---------------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('case', COMBINER, ids=[c['id'] for c in COMBINER])
def test_background_compose_matches_reference(case):
    combiner, _ = _combiner_for(case)
    height, width = case['canvas']
    for run in case['runs']:
        rng = np.random.default_rng(run['rng_seed'])
        image = combiner.run(height, width, rng)
        assert image.on_device and image.shape == (height, width)
        got = image.mat
        ref = f4_array(case, 'image') if run is case['runs'][0] else None
        if ref is not None:
            assert np.array_equal(got, ref), np.argwhere((got != ref).any(axis=2))[:5]
        assert sha(got) == run['sha'], (case['id'], run['rng_seed'])
        assert int(rng.integers(0, 2**31)) == run['rng_after']


@pytest.mark.gpu
@pytest.mark.parametrize('case', TEXT_LINES, ids=[c['id'] for c in TEXT_LINES])
def test_atlas_text_line_matches_reference(case):
    from vkit_b200.compositing import GlyphAtlas, render_atlas_glyphs_in_text_line
    from vkit_b200.element import Box
    atlas = GlyphAtlas(page_bytes=1 << 16)
    bitmaps = f4_glyph_bitmaps(case['seed'], case['count'], case['lcd'])
    glyphs = []
    half = len(bitmaps) // 2
    for j, (bitmap, top, left, advance_x) in enumerate(bitmaps):
        glyphs.append(atlas.add(('font', j), bitmap, gamma=case['gamma'], bitmap_top=top,
                                bitmap_left=left, advance_x=advance_x))
        if j == half:
            atlas.commit()  # two commits: the second chunk must not disturb the first
    atlas.commit()
    assert atlas.add(('font', 0), bitmaps[0][0]) is glyphs[0]  # cache hit, nothing pending
    for glyph, metrics in zip(glyphs, case['metrics']):
        assert sha(glyph.bitmap.to_host()) == metrics['sha_image']
        if not case['lcd']:
            assert sha(glyph.alpha.to_host()) == metrics['sha_alpha']
    boxes = [Box(up=b[0], down=b[1], left=b[2], right=b[3]) for b in case['boxes']]
    height, width = case['line_shape']
    image, mask, score_map = render_atlas_glyphs_in_text_line(tuple(case['glyph_color']), height,
                                                              width, glyphs, boxes)
    ref_image = f4_array(case, 'image')
    assert np.array_equal(image.mat, ref_image)
    assert sha(image.mat) == case['sha']['image'] and sha(mask.mat) == case['sha']['mask']
    if case['lcd']:
        assert score_map is None
    else:
        assert sha(score_map.mat) == case['sha']['score_map']


@pytest.mark.gpu
def test_lcd_text_line_from_host_arrays():
    """The host-array entry point takes H x W x 3 glyph images (freetype.py:355-370)."""
    from oracle import vkit_port as port
    from vkit_b200.compositing import render_char_glyphs_in_text_line
    from vkit_b200.element import Box
    case = next(c for c in TEXT_LINES if c['lcd'])
    planes = [port.glyph_planes(b[0], case['gamma'])
              for b in f4_glyph_bitmaps(case['seed'], case['count'], True)]
    boxes = [Box(up=b[0], down=b[1], left=b[2], right=b[3]) for b in case['boxes']]
    height, width = case['line_shape']
    image, mask, score_map = render_char_glyphs_in_text_line(
        tuple(case['glyph_color']), height, width, [p[0] for p in planes], None, boxes,
        glyph_color_gamma=case['gamma'])
    assert score_map is None
    assert sha(image.mat) == case['sha']['image'] and sha(mask.mat) == case['sha']['mask']


@pytest.mark.gpu
@pytest.mark.parametrize('lcd', [False, True])
def test_batched_text_lines_match_reference(lcd):
    """All text lines of a page through three launches (render_atlas_text_lines): every line equals
    the live-reference fixture of its case."""
    from vkit_b200.compositing import GlyphAtlas, render_atlas_text_lines
    from vkit_b200.element import Box
    cases = [c for c in TEXT_LINES if c['lcd'] == lcd]
    atlas = GlyphAtlas(page_bytes=1 << 18)
    lines = []
    for case in cases:
        bitmaps = f4_glyph_bitmaps(case['seed'], case['count'], case['lcd'])
        glyphs = [atlas.add((case['id'], j), bitmap, gamma=case['gamma'], bitmap_top=top,
                            bitmap_left=left, advance_x=advance_x)
                  for j, (bitmap, top, left, advance_x) in enumerate(bitmaps)]
        boxes = [Box(up=b[0], down=b[1], left=b[2], right=b[3]) for b in case['boxes']]
        if case is cases[-1]:  # the array form of the boxes: (up, left) per glyph
            boxes = np.asarray([(b[0], b[2]) for b in case['boxes']])
        height, width = case['line_shape']
        lines.append((tuple(case['glyph_color']), height, width, glyphs, boxes))
    atlas.commit()
    rendered = render_atlas_text_lines(lines)
    assert len(rendered) == len(cases)
    for case, (image, mask, score_map) in zip(cases, rendered):
        assert sha(image.mat) == case['sha']['image'], case['id']
        assert sha(mask.mat) == case['sha']['mask'], case['id']
        if lcd:
            assert score_map is None
        else:
            assert sha(score_map.mat) == case['sha']['score_map'], case['id']


def test_textures_from_folder_match_in_memory_textures(tmp_path):
    """`<folder>/metas.json` + `<folder>/image/*` (the layout ImageCombinerEngine reads,
    combiner.py:48-69): same pixels, same walk as textures handed over in memory."""
    import json
    from PIL import Image as PilImage
    from vkit_b200.background import ImageCombiner, load_textures_from_folder
    case = COMBINER[0]
    textures = f4_textures(case['textures_seed'], case['textures_count'], *case['size_range'])
    (tmp_path / 'image').mkdir()
    metas = []
    for name, mat, mean, std in textures:
        PilImage.fromarray(mat).save(tmp_path / 'image' / name)
        metas.append({'image_file': name, 'grayscale_mean': mean, 'grayscale_std': std})
    (tmp_path / 'metas.json').write_text(json.dumps(metas))
    loaded = load_textures_from_folder(str(tmp_path))
    assert len(loaded) == len(textures)
    for texture, (name, mat, mean, std) in zip(loaded, textures):
        assert texture.name.endswith(name) and np.array_equal(texture.image.mat, mat)
        assert texture.grayscale_mean == mean and texture.grayscale_std == std
    from_folder = ImageCombiner(loaded)
    in_memory, _ = _combiner_for(case)
    height, width = case['canvas']
    run = case['runs'][0]
    plans = []
    for combiner in (from_folder, in_memory):
        rng = np.random.default_rng(run['rng_seed'])
        placements = combiner.plan(height, width, combiner.sample_candidates(rng), rng)
        plans.append([[p.up, p.down, p.left, p.right] for p in placements])
    assert plans[0] == plans[1] == run['rects']

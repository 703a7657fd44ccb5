"""Host side of DecisionTreeDatasetConfig (src/decision_tree.py:21-122) without a GPU: the descriptor form (num_images = 0, as
test_on_saved_model.py uses it for the colour table) and the colour <-> id conversions against the reference's per-class loops
restated naively."""
import json

import numpy as np
import pytest

from rdf_b200.decision_tree import DecisionTreeDatasetConfig


@pytest.fixture
def descriptor(tmp_path):
    colors = {'1': [255, 0, 0, 255], '2': [0, 255, 0, 255], '3': [0, 0, 255, 255], '7': [12, 34, 56, 78]}
    (tmp_path / 'config.json').write_text(json.dumps({'img_dims': [24, 10], 'num_images': 5, 'id_to_color': colors}))
    return DecisionTreeDatasetConfig(str(tmp_path) + '/')


def test_descriptor_only_loads_nothing(descriptor):
    assert descriptor.num_images == 0 and descriptor.total_available_images == 5
    assert descriptor.img_dims == (24, 10) and descriptor.num_classes() == 5          # background id 0 counts
    assert not hasattr(descriptor, 'depth_blocks')
    assert descriptor.images_shape() == (0, 10, 24) and descriptor.num_pixels() == 0
    assert descriptor.id_to_color[0].tolist() == [0, 0, 0, 0] and descriptor.id_to_color[7].dtype == np.uint8


def test_colour_conversions_match_the_per_class_loops(descriptor):
    rng = np.random.default_rng(3)
    known = np.array(sorted(descriptor.id_to_color))
    ids = known[rng.integers(0, len(known), size=(3, 10, 24))].astype(np.uint16)
    ids[0, 0, :4] = [4, 5, 6, 300]                                                   # ids without a colour
    colours = descriptor.convert_ids_to_colors(ids)
    want = np.zeros(ids.shape + (4,), dtype=np.uint8)                                 # src/decision_tree.py:101-110
    for class_id, color in descriptor.id_to_color.items():
        want[np.where(ids == class_id)] = color
    assert colours.dtype == np.uint8 and np.array_equal(colours, want)

    picture = colours[1]                                                              # every pixel carries a known colour
    back = descriptor.convert_colors_to_ids(picture)
    assert back.dtype == np.uint16 and back.shape == (10, 24) and np.array_equal(back, ids[1])
    with pytest.raises(AssertionError):                                               # src/decision_tree.py:97: unknown colour
        bad = picture.copy()
        bad[2, 3] = [1, 2, 3, 4]
        descriptor.convert_colors_to_ids(bad)


def test_empty_id_image_converts(descriptor):
    assert descriptor.convert_ids_to_colors(np.zeros((0, 10, 24), dtype=np.uint16)).shape == (0, 10, 24, 4)

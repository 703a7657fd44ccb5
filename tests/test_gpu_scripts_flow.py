"""The call sequences of the reference's drop-in target scripts, executed through the compat shim on a synthetic on-disk dataset:
  * train_model.py:70-139   (DecisionTreeDatasetConfig.multiple -> trainer.allocate/train -> get_labels -> forest -> np.save)
  * test_on_saved_model.py:38-67   (DecisionForest.load -> dataset block fetch -> get_labels_forest -> renders)
  * run_live_layered.py:40,126 / 3d_bz.py:437,461   (LayeredDecisionForest.load(json) -> run -> MeanShift.run)
Only the window / camera / GL parts of the scripts are left out (out of scope)."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_dataset(path, num, W, H, seed=0):
    """dataset directory in the reference's format (src/decision_tree.py:46-83, writer src/live_data_convert.py:287-298,448,458):
    config.json + %08d_depth.png / %08d_labels.png (16-bit).  A near square (label 2) on a far background (label 1)."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    os.makedirs(path, exist_ok=True)
    for i in range(num):
        depth = (4000 + rng.integers(0, 16, size=(H, W))).astype(np.uint16)
        labels = np.ones((H, W), np.uint16)
        y0, x0 = rng.integers(4, H - 44), rng.integers(4, W - 44)
        depth[y0:y0 + 40, x0:x0 + 40] = (2500 + rng.integers(0, 16, size=(40, 40))).astype(np.uint16)
        labels[y0:y0 + 40, x0:x0 + 40] = 2
        depth[0:2, :] = 0                                        # missing pixels: label 0, never trained or scored
        labels[0:2, :] = 0
        Image.fromarray(depth).save(os.path.join(path, f'{i:08d}_depth.png'))
        Image.fromarray(labels).save(os.path.join(path, f'{i:08d}_labels.png'))
    cfg = {'img_dims': [W, H], 'num_images': num, 'id_to_color': {'1': [255, 0, 0, 255], '2': [0, 255, 0, 255]}}
    with open(os.path.join(path, 'config.json'), 'w') as f:
        json.dump(cfg, f)


def test_train_model_and_test_on_saved_model_flows(tmp_path):
    compat = os.path.join(ROOT, '3d-beats_b200', 'compat')
    if compat not in sys.path:
        sys.path.insert(0, compat)
    import rdf_dropin  # noqa: F401
    ns = {}
    exec('from decision_tree import *', ns)                      # the scripts' own import line
    DecisionTreeTrainer, DecisionTreeEvaluator = ns['DecisionTreeTrainer'], ns['DecisionTreeEvaluator']
    DecisionTreeDatasetConfig, DecisionTree, DecisionForest = ns['DecisionTreeDatasetConfig'], ns['DecisionTree'], ns['DecisionForest']
    cu_array, MAX_UINT16 = ns['cu_array'], ns['MAX_UINT16']

    W, H = 160, 120
    data = str(tmp_path / 'data') + '/'
    _write_dataset(data, 12, W, H)
    np.random.seed(3)                                            # proposals come from np.random, as in the reference

    # ---- train_model.py:70-139 ----
    NUM_TRAIN, NUM_TEST, F, P, D, TREES = 8, 4, 128, 64, 6, 2
    trainer = DecisionTreeTrainer(NUM_TRAIN, P)
    evaluator = DecisionTreeEvaluator()
    train_data, test_data = DecisionTreeDatasetConfig.multiple(data, [(NUM_TRAIN, NUM_TRAIN, 'train'), (NUM_TEST, None, 'test')])
    assert train_data.num_classes() == 3 and test_data.images_shape() == (NUM_TEST, H, W)
    tree1 = DecisionTree(D, train_data.num_classes())
    trainer.allocate(train_data, F, tree1.max_depth)
    test_output_labels_cu = cu_array.GPUArray(test_data.images_shape(), dtype=np.uint16)
    tree_cpu = np.zeros((tree1.TOTAL_TREE_NODES, tree1.TREE_NODE_ELS), dtype=np.float32)
    forest_cpu = np.zeros((TREES, tree1.TOTAL_TREE_NODES, tree1.TREE_NODE_ELS), dtype=np.float32)
    test_labels_cu = cu_array.GPUArray(test_data.images_shape(), dtype=np.uint16)
    test_depth_cu = cu_array.GPUArray(test_data.images_shape(), dtype=np.uint16)
    for i in range(TREES):
        test_data.get_depth_block_cu(0, test_depth_cu)
        test_data.get_labels_block_cu(0, test_labels_cu)
        test_labels_cpu = test_labels_cu.get()
        trainer.train(train_data, tree1)
        test_output_labels_cu.fill(MAX_UINT16)
        evaluator.get_labels(tree1, test_depth_cu, test_output_labels_cu)
        test_output_labels = test_output_labels_cu.get()
        pct_match = np.sum(test_output_labels == test_labels_cpu) / np.sum(test_labels_cpu > 0)
        assert pct_match > 0.9, pct_match                         # a depth-6 tree separates a near square from a far wall
        tree1.tree_out_cu.get(tree_cpu)
        forest_cpu[i] = np.copy(tree_cpu)
    forest1 = DecisionForest(TREES, D, test_data.num_classes())
    forest1.forest_cu.set(forest_cpu)
    test_output_labels_cu.fill(np.uint16(MAX_UINT16))
    evaluator.get_labels_forest(forest1, test_depth_cu, test_output_labels_cu)
    forest_out = test_output_labels_cu.get()
    pct_forest = np.sum(forest_out == test_labels_cpu) / np.sum(test_labels_cpu > 0)
    assert pct_forest > 0.9
    assert (forest_out[test_labels_cpu == 0] == 65535).all()      # missing pixels are not classified
    model = str(tmp_path / 'model.npy')
    np.save(model, forest_cpu)

    # the trained forest evaluates identically under the CPU oracle
    from oracle import c_oracle as co
    exp = np.full(test_data.images_shape(), 65535, np.uint16)
    co.eval_forest(forest_cpu, test_depth_cu.get(), exp)
    assert np.array_equal(forest_out, exp)

    # ---- test_on_saved_model.py:38-67 ----
    forest = DecisionForest.load(model)
    assert (forest.num_trees, forest.max_depth, forest.num_classes) == (TREES, D, 3)
    dataset = DecisionTreeDatasetConfig(data, num_images=NUM_TEST, imgs_name='test')
    d_cu = cu_array.GPUArray(dataset.images_shape(), dtype=np.uint16)
    dataset.get_depth_block_cu(0, d_cu)
    l_cu = cu_array.GPUArray(dataset.images_shape(), dtype=np.uint16)
    dataset.get_labels_block_cu(0, l_cu)
    out_cu = cu_array.GPUArray(dataset.images_shape(), dtype=np.uint16)
    out_cu.fill(MAX_UINT16)
    evaluator.get_labels_forest(forest, d_cu, out_cu)
    out = out_cu.get()
    labels_cpu = l_cu.get()
    assert np.sum(out == labels_cpu) / np.sum(labels_cpu > 0) > 0.85
    render = dataset.convert_ids_to_colors(np.where(out == 65535, 0, out))
    assert render.shape == (NUM_TEST, H, W, 4) and render.dtype == np.uint8


def test_layered_json_model_and_mean_shift_flow(tmp_path):
    """run_live_layered.py:40,126 and 3d_bz.py:437,461-465: layered model from JSON + .npy files, run(), MeanShift.run()."""
    import torch
    from rdf_b200 import synth
    from rdf_b200 import decision_tree as dt
    from rdf_b200.buffers import GpuBuffer
    from rdf_b200.mean_shift import MeanShift
    from oracle import numpy_oracle as no
    H, W, r = 240, 424, 2
    forests, cfg, variances = synth.layered_cfg2(max_depth=9)
    path = synth.write_layered_model(str(tmp_path / 'model'), forests, cfg)
    ldf = dt.LayeredDecisionForest.load(path, (H, W), labels_reduce=r)
    assert ldf.num_layered_classes == 11 and ldf.labels_dims == (H // r, W // r) and len(ldf.label_images) == 2
    depth = synth.depth_frames('live-mask', 1, H, W, seed=77)
    depth_image = GpuBuffer((H, W), np.uint16)
    depth_image.cu().set(depth[0])
    labels_image = GpuBuffer((1, H // r, W // r), np.uint16)
    ldf.run(depth_image, labels_image, W / 848.)                               # scale_factor = DIM_X / 848 (run_live_layered.py:126)
    means = MeanShift().run(6, labels_image.cu(), ldf.num_layered_classes, variances)
    torch.cuda.synchronize()
    comp, _ = no.layered_run(forests, [(None, None), (0, 1)], cfg['conditions'], depth[0], r, W / 848.)
    assert np.array_equal(labels_image.cu().get()[0], comp)
    exp = no.mean_shift(comp, 11, variances, 6)
    assert means.shape == (11, 2) and np.array_equal(np.isnan(means), np.isnan(exp)) and np.nanmax(np.abs(means - exp)) <= 1e-5

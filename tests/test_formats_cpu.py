"""On-disk formats around the box pipeline against tests/golden/formats.npz, produced by the reference's
own AnnotationTransform and by verbatim restatements of its PR-data scripts (oracle/make_golden.py gen_formats).  CPU only."""
import numpy as np

from fdt_b200.utils import formats


def test_annotation_lines(golden, tmp_path):
    g = golden("formats")
    lines = [str(s) for s in g["lines"]]
    p = tmp_path / "anno"
    p.write_text("\n".join(lines) + "\n")
    ids, ann = formats.read_annotation_file(str(p))
    assert ids == [ln.split()[0] for ln in lines]
    tr = formats.AnnotationTransform()
    for i, a in enumerate(ann):
        got = np.array(tr(a, 640, 480), dtype=np.float64).reshape(-1, 5)
        assert np.array_equal(got, g[f"at_{i}"])
        assert np.array_equal(formats.annotation_to_pixel_boxes(a), g[f"px_{i}"])
    t = formats.annotation_to_targets(ann, [(640, 480)] * len(ann))
    assert [tuple(x.shape) for x in t] == [g[f"at_{i}"].shape for i in range(len(ann))]
    assert np.array_equal(t[0].numpy(), g["at_0"].astype(np.float32))


def test_pr_data_roundtrip(golden, tmp_path):
    g = golden("formats")
    acc = formats.new_pr_accumulator()
    acc = formats.accumulate_pr(acc, g["pr_tf_conf"])
    path = str(tmp_path / "data_of_repo.npy")
    data = formats.save_pr_data(path, acc, int(g["pr_truth_num"]))
    assert np.array_equal(data, g["pr_file"]) and np.array_equal(np.load(path), g["pr_file"])
    tf_conf, truth_num = formats.load_pr_data(path)
    assert truth_num == g["pr_truth_num"]
    tp, fp = formats.gen_tp_fp(tf_conf)
    assert np.array_equal(tp, g["pr_tp"]) and np.array_equal(fp, g["pr_fp"])
    (recall, precision), (fp2, recall2) = formats.pr_roc(np.load(path))
    assert np.array_equal(recall, g["pr_recall"]) and np.array_equal(precision, g["pr_precision"])
    assert np.array_equal(fp2, fp) and np.array_equal(recall2, recall)

"""CPU tests of the eval-harness marshalling (reference tools/cocotools.py:159-191)."""
import numpy as np


def reference_records(im_id, boxes, scores, classes, clsid2catid):
    """The reference's per-detection loop (tools/cocotools.py:166-186), restated verbatim in spirit."""
    out = []
    for p in range(len(boxes)):
        xmin, ymin, xmax, ymax = boxes[p]
        w = xmax - xmin + 1
        h = ymax - ymin + 1
        bbox = [round(float(x) * 10) / 10 for x in [xmin, ymin, w, h]]
        out.append({'image_id': im_id, 'category_id': clsid2catid[int(classes[p])], 'bbox': bbox, 'score': float(scores[p])})
    return out


def test_detections_to_coco_matches_reference_loop():
    from tools.cocotools import detections_to_coco, clsid2catid, catid2clsid
    assert len(clsid2catid) == 80 and clsid2catid[0] == 1 and clsid2catid[79] == 90 and catid2clsid[13] == 11
    rs = np.random.RandomState(0)
    boxes = (rs.rand(300, 4) * 600).astype(np.float32)
    boxes[:, 2:] += boxes[:, :2]
    boxes[:7] = np.array([[0.05, 0.15, 10.25, 20.35]] * 7, dtype=np.float32) + np.arange(7, dtype=np.float32)[:, None] * 0.1   # .x5 ties
    scores = rs.rand(300).astype(np.float32)
    classes = rs.randint(0, 80, 300).astype(np.int32)
    got = detections_to_coco(42, boxes, scores, classes, clsid2catid)
    assert got == reference_records(42, boxes, scores, classes, clsid2catid)
    assert detections_to_coco(1, np.array([]), np.array([]), np.array([]), clsid2catid) == []


def test_eval_harness_end_to_end_with_fake_decode(tmp_path, monkeypatch):
    """The harness loop (reader thread, batches of any size, one merged json) with a stand-in Decode: every image must be
    visited exactly once, in order, and the records must follow the detections handed back."""
    import json
    import cv2
    from tools import cocotools
    monkeypatch.chdir(tmp_path)
    os_dir = tmp_path / 'imgs'
    os_dir.mkdir()
    images = []
    for i in range(7):
        name = 'im%03d.jpg' % i
        cv2.imwrite(str(os_dir / name), np.full((20 + i, 30, 3), i * 10, dtype=np.uint8))
        images.append({'id': 100 + i, 'file_name': name})

    class FakeDecode(object):
        use_gpu = False

        def process_image(self, img):
            return np.zeros((1, 3, 8, 8), dtype=np.float32), np.array([[img.shape[0], img.shape[1]]], dtype=np.int32)

        def detect_batch(self, batch_img, batch_pimage, batch_im_size, draw_image, draw_thresh=0.0):
            n = len(batch_img)
            assert batch_pimage.shape[0] == n and batch_im_size.shape == (n, 2)
            boxes = [np.array([[1.0, 2.0, float(s[1]), float(s[0])]], dtype=np.float32) for s in batch_im_size]
            return batch_img, boxes, [np.array([0.5], dtype=np.float32)] * n, [np.array([3], dtype=np.int32)] * n

    recs = cocotools.eval(FakeDecode(), images, str(os_dir) + '/', None, 3, cocotools.clsid2catid, False, 0.0, type='test_dev')
    assert [r['image_id'] for r in recs] == [100 + i for i in range(7)]
    assert recs[2]['bbox'] == [1.0, 2.0, 30.0, 21.0] and recs[2]['category_id'] == 4
    assert json.load(open('results/bbox_detections.json')) == recs

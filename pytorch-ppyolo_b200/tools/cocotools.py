"""COCO evaluation harness around the detector (reference tools/cocotools.py:30-277), same entry points --
``eval(_decode, images, eval_pre_path, anno_file, eval_batch_size, _clsid2catid, draw_image, draw_thresh, type)``,
``bbox_eval``, ``cocoapi_eval``, ``get_classes``, ``catid2clsid`` / ``clsid2catid`` -- restructured for a detector that runs
at thousands of images per second:

* the reader prefetches whole batches (bounded queue of 3, like the reference's dict polling, tools/cocotools.py:111-157) and
  hands the engine the RESIZED uint8 images (``Decode.process_image_u8``): normalisation and layout change happen inside
  the first kernel, a quarter of the upload bytes;
* detections of a batch are marshalled to COCO records in ONE vectorised pass (``detections_to_coco``: the reference's
  ``w = xmax - xmin + 1`` convention and round-to-0.1, tools/cocotools.py:174-179, as numpy array ops) instead of one thread and
  one json file per image that are read back afterwards (:159-191, :84-92);
* ``eval_batch_size`` is whatever the caller passes (the reference config's 4 works; 32+ is what the engine wants).

pycocotools is imported lazily by ``cocoapi_eval`` exactly like the reference; without it ``eval`` still writes
``eval_results/bbox_detections.json`` and returns the records."""
import json
import logging
import os
import queue
import threading
import time

import numpy as np

logger = logging.getLogger(__name__)

_COCO_IDS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 27, 28, 31, 32, 33, 34, 35, 36, 37, 38, 39,
             40, 41, 42, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 67, 70, 72, 73, 74, 75, 76, 77,
             78, 79, 80, 81, 82, 84, 85, 86, 87, 88, 89, 90]
clsid2catid = {i: c for i, c in enumerate(_COCO_IDS)}          # reference tools/cocotools.py:22-28
catid2clsid = {c: i for i, c in enumerate(_COCO_IDS)}          # :30-36


def get_classes(classes_path):
    with open(classes_path) as f:
        return [c.strip() for c in f.readlines()]


def detections_to_coco(im_id, boxes, scores, classes, _clsid2catid):
    """One image's detections -> list of COCO result records (reference multi_thread_write_json, tools/cocotools.py:159-186):
    bbox = [xmin, ymin, xmax - xmin + 1, ymax - ymin + 1], every number rounded to 0.1 with Python's round (half to even)."""
    if boxes is None or len(boxes) == 0:
        return []
    b = np.asarray(boxes, dtype=np.float32)
    x0, y0, x1, y1 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    bbox = np.stack([x0, y0, x1 - x0 + 1, y1 - y0 + 1], axis=1)             # float32 arithmetic, like the reference's numpy scalars
    bbox = np.round(bbox.astype(np.float64) * 10) / 10                        # round(float(x) * 10) / 10
    cats = [_clsid2catid[int(c)] for c in classes]
    sc = np.asarray(scores, dtype=np.float32)
    return [{'image_id': im_id, 'category_id': cats[i], 'bbox': [float(v) for v in bbox[i]], 'score': float(sc[i])}
            for i in range(len(cats))]


def cocoapi_eval(jsonfile, style, coco_gt=None, anno_file=None, max_dets=(100, 300, 1000)):
    """Reference tools/cocotools.py:44-75."""
    assert coco_gt is not None or anno_file is not None
    from pycocotools.coco import COCO
    from pycocotools.cocoeval import COCOeval
    if coco_gt is None:
        coco_gt = COCO(anno_file)
    logger.info('Start evaluate...')
    coco_dt = coco_gt.loadRes(jsonfile)
    if style == 'proposal':
        coco_eval = COCOeval(coco_gt, coco_dt, 'bbox')
        coco_eval.params.useCats = 0
        coco_eval.params.maxDets = list(max_dets)
    else:
        coco_eval = COCOeval(coco_gt, coco_dt, style)
    coco_eval.evaluate()
    coco_eval.accumulate()
    coco_eval.summarize()
    return coco_eval.stats


def bbox_eval(anno_file, outfile='eval_results/bbox_detections.json'):
    """Reference tools/cocotools.py:77-98 (the merged json is already written by ``eval``)."""
    from pycocotools.coco import COCO
    return cocoapi_eval(outfile, 'bbox', coco_gt=COCO(anno_file))


def _read_batches(images, _decode, eval_pre_path, eval_batch_size, out_q, use_u8):
    import cv2
    n = len(images)
    for start in range(0, n, eval_batch_size):
        chunk = images[start:start + eval_batch_size]
        slots = [None] * len(chunk)

        def load(j, im):
            image = cv2.imread(eval_pre_path + im['file_name'])
            pimage, im_size = (_decode.process_image_u8 if use_u8 else _decode.process_image)(np.copy(image))
            slots[j] = (im['id'], im['file_name'], image, pimage, im_size)
        threads = [threading.Thread(target=load, args=(j, im)) for j, im in enumerate(chunk)]     # cv2 releases the GIL
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        out_q.put({'batch_im_id': [s[0] for s in slots], 'batch_im_name': [s[1] for s in slots], 'batch_img': [s[2] for s in slots],
                   'batch_pimage': np.concatenate([s[3] for s in slots], axis=0), 'batch_im_size': np.concatenate([s[4] for s in slots], axis=0)})
    out_q.put(None)


def eval(_decode, images, eval_pre_path, anno_file, eval_batch_size, _clsid2catid, draw_image, draw_thresh, type='eval', use_u8=True):
    """Reference tools/cocotools.py:195-277: run the detector over ``images`` (COCO image records), write the detections in COCO
    result format and (type='eval') score them with pycocotools.  Returns the COCOeval stats, or the records when pycocotools is
    not installed / for 'test_dev'."""
    assert type in ['eval', 'test_dev']
    result_dir = 'eval_results' if type == 'eval' else 'results'
    os.makedirs(result_dir, exist_ok=True)
    if draw_image:
        os.makedirs('%s/images' % result_dir, exist_ok=True)
    n = len(images)
    q = queue.Queue(maxsize=3)
    reader = threading.Thread(target=_read_batches, args=(images, _decode, eval_pre_path, eval_batch_size, q, use_u8 and _decode.use_gpu), daemon=True)
    start = time.time()
    reader.start()
    records, it = [], 0
    while True:
        dic = q.get()
        if dic is None:
            break
        imgs, boxes, scores, classes = _decode.detect_batch(dic['batch_img'], dic['batch_pimage'], dic['batch_im_size'], draw_image=draw_image,
                                                            draw_thresh=draw_thresh)
        for j in range(len(boxes)):
            records += detections_to_coco(dic['batch_im_id'][j], boxes[j], scores[j], classes[j], _clsid2catid)
            if draw_image:
                import cv2
                cv2.imwrite('%s/images/%s' % (result_dir, dic['batch_im_name'][j]), imgs[j])
        if it % 100 == 0:
            logger.info('Test iter {}'.format(it))
        it += 1
    cost = time.time() - start
    logger.info('total time: {0:.6f}s'.format(cost))
    logger.info('Speed: %.6fs per image,  %.1f FPS.' % (cost / max(n, 1), n / max(cost, 1e-9)))
    outfile = '%s/bbox_detections.json' % result_dir
    with open(outfile, 'w') as f:
        json.dump(records, f)
    if type == 'eval':
        try:
            return bbox_eval(anno_file, outfile)
        except ImportError:
            logger.warning('pycocotools is not installed: detections written to %s, not scored', outfile)
    return records

"""``GeneralizedRCNN`` (inference) over the B200 plug-ins.

Call sequence of detectron2's ``GeneralizedRCNN.inference`` [d2] as mirrored in-tree by the reference
at ``/root/reference/tester.py:24-75``; the tensor-in / tuple-out variant follows
``/root/reference/modified_class.py:27-40`` and ``deploy_utils.py:117-126``.
"""
import collections

import torch
from torch import nn

from .. import runtime
from .compat import (BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY, Boxes,
                     ImageList, Instances, ShapeSpec)
from .fcos import instances_from_det


class _Sizes(object):
    """What FCOS / ROI heads read from an ImageList: ``len()`` and ``.image_sizes``."""

    def __init__(self, image_sizes):
        self.image_sizes = [tuple(s) for s in image_sizes]

    def __len__(self):
        return len(self.image_sizes)


class GeneralizedRCNN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.backbone = BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, ShapeSpec(channels=len(cfg.MODEL.PIXEL_MEAN)))
        shape = self.backbone.output_shape()
        self.proposal_generator = PROPOSAL_GENERATOR_REGISTRY.get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(cfg, shape)
        self.roi_heads = ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, shape)
        self.register_buffer("pixel_mean", torch.tensor(cfg.MODEL.PIXEL_MEAN).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(cfg.MODEL.PIXEL_STD).view(-1, 1, 1), False)
        self.train(False)

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("inference only")
        return super().train(False)

    @property
    def device(self):
        return runtime.engine_for(self.cfg).device

    # -- the reference-facing call -------------------------------------------------------------
    def forward(self, batched_inputs):
        return self.inference(batched_inputs)

    @torch.no_grad()
    def inference_stream(self, batches, do_postprocess=True, depth=2, with_event=False):
        """Generator over an iterable of ``batched_inputs`` lists: yields what ``inference`` returns for each, in order.

        Same results as calling ``forward`` per batch, software-pipelined ``depth`` batches deep: (1) the host->device
        copy of the next batch (pinned host images) runs on a copy stream into a staging buffer while earlier batches
        are computed; (2) the device work of the following ``depth`` batches is enqueued BEFORE the host waits for batch
        ``i`` and cuts its ``Instances`` out, so the GPU does not idle during the host-side assembly and read-back.
        Every batch is still copied from the host, run and read back; only the order of issue changes.

        ``with_event=True`` yields ``(results, event)``: the CUDA event marks the end of that batch's device work.  A
        consumer that reads results back should do so on its own stream after ``stream.wait_event(event)`` -- a copy on
        the compute stream would queue behind the batches already enqueued ahead and drain the pipeline."""
        eng = runtime.engine_for(self.cfg)
        it = iter(batches)
        stage, consume = self._stager(eng)

        pending = collections.deque()
        state = {"staged": None, "launched": 0}
        state["staged"] = stage(next(it, None))

        def launch_next():
            # device side of the next batch: staged pixels -> input buffers, following batch's H2D, then the whole step
            if state["staged"] is None:
                return False
            consume(state["staged"])
            batch = state["staged"][0]
            state["staged"] = stage(next(it, None))
            pending.append(self._launch(batch, do_postprocess, True, state["launched"] % (depth + 1)))
            state["launched"] += 1
            return True

        for _ in range(depth):
            if not launch_next():
                break
        while pending:
            launch_next()                               # keep `depth` batches queued behind the one about to be finished
            ctx = pending.popleft()
            out = self._finish(ctx)
            yield (out, ctx["done"]) if with_event else out

    @staticmethod
    def _stager(eng):
        """(stage, consume) of the pipelined entry points: ``stage(batch)`` issues the host->device copies of a batch of
        pinned host images on the copy stream into staging buffers; ``consume(staged)`` makes the compute stream wait for
        them and moves the pixels into the input buffers the launch plan reads (device to device, ~0.03 ms for 51 MB)."""
        copy_stream = eng.copy_stream()

        def stage(batch):
            if batch is None:
                return None
            if eng._stage_free is not None:                  # the previous consumer of the staging buffers has read them
                copy_stream.wait_event(eng._stage_free)
            with torch.cuda.stream(copy_stream):
                sig = tuple((tuple(b["image"].shape), b["image"].dtype) for b in batch)
                bufs, whole = eng.image_buffers("stage_image", sig)
                for dst, b in zip(bufs, batch):
                    dst.copy_(b["image"], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return batch, bufs, ev, whole, sig

        def consume(staged):
            _, bufs, ev, whole, sig = staged
            torch.cuda.current_stream().wait_event(ev)
            images, whole_in = eng.image_buffers("input_image", sig)
            if whole is not None and whole_in is not None:
                whole_in.copy_(whole, non_blocking=True)                # one copy for the batch (images of one shape)
            else:
                for dst, src in zip(images, bufs):
                    dst.copy_(src, non_blocking=True)
            eng._stage_free = torch.cuda.Event()
            eng._stage_free.record()

        return stage, consume

    @torch.no_grad()
    def inference_records(self, batches, depth=2, rle_capacity=1 << 22):
        """Pipelined inference that hands the WHOLE result of every batch to the host: generator over an iterable of
        ``batched_inputs`` lists (pinned host images), yielding one ``parallel.BatchResult`` per batch, in order.

        What the reference returns per image is ``Instances`` with boxes, scores, classes, locations, mask scores and
        full-resolution bool masks (1 MB per mask at 800x1333); what leaves the device here is the fixed-size detection
        record of every slot (``parallel.RECORD_FIELDS``: post-processed box, score, class, mask score, location, validity)
        and the masks as COCO run lengths (``csrc/rle.cu``: pycocotools' column-major ``rleEncode``, the form
        ``instances_to_coco_json`` -- coco_evaluation.py:362-427 -- turns them into anyway), computed on the device right
        after the paste-back.  Per batch: H2D of the images (copy stream, overlapping the previous batch), one graph
        replay, paste-back, RLE, then two small D2H copies on a read-back stream (records + run offsets, then exactly the
        runs that were produced).  ``depth`` batches are enqueued ahead of the one being read back, so the device never
        waits for the host.  ``rle_capacity``: initial run capacity of the device / pinned run buffers (they grow when a
        batch needs more: that batch is re-encoded, nothing is lost)."""
        from .. import lib, parallel
        eng = runtime.engine_for(self.cfg)
        it = iter(batches)
        copy_stream = eng.copy_stream()
        back = eng.side_stream("readback")
        fcos, roi = self.proposal_generator, self.roi_heads
        nslots = depth + 1
        state = {"staged": None, "launched": 0, "cap": int(rle_capacity)}
        pending = collections.deque()
        slot_free = eng.records_slot_free                 # slot -> event: the step that last used the slot has finished

        # A batch lives in one of ``depth + 1`` SLOTS: its own input buffers (the H2D target), mask / run / record buffers.
        # The WHOLE step of a slot -- fused stem straight from the slot's input buffers ... paste-back, RLE, record packing --
        # is one CUDA graph, so the compute stream sees [wait for the H2D event][graph][event] per batch and nothing else
        # (profiles/r2_trace_e2e_records_b16.txt: with eager staging copies / paste / RLE launches between the graphs the
        # device idled 0.2-0.4 ms per step at those hand-offs).
        def stage(batch, slot):
            if batch is None:
                return None
            sig = tuple((tuple(b["image"].shape), b["image"].dtype) for b in batch)
            if slot in slot_free:                            # the step that last used this slot has finished with the inputs
                copy_stream.wait_event(slot_free[slot])
            with torch.cuda.stream(copy_stream):
                bufs, _ = eng.image_buffers("records_in{}_".format(slot), sig)
                for dst, b in zip(bufs, batch):
                    dst.copy_(b["image"], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return batch, bufs, ev, sig, slot

        def launch_next():
            if state["staged"] is None:
                return False
            batch, images, ev, sig, slot = state["staged"]
            # the H2D copy of this batch was issued a whole step ago: waiting for it on the HOST returns at once in steady state
            # and keeps the compute stream free of cross-stream waits (a wait in front of a graph launch exposed ~0.1 ms of
            # launch latency per step)
            ev.synchronize()
            state["launched"] += 1
            state["staged"] = stage(next(it, None), (slot + 1) % nslots)
            n = len(batch)
            sizes = [(int(b["image"].shape[-2]), int(b["image"].shape[-1])) for b in batch]
            out_sizes = [(int(b.get("height", sz[0])), int(b.get("width", sz[1]))) for b, sz in zip(batch, sizes)]
            if len(set(out_sizes)) != 1:
                raise ValueError("inference_records needs one output size per batch (got {})".format(sorted(set(out_sizes))))
            oh, ow = out_sizes[0]
            cap = state["cap"]
            B = eng.buffer

            def plan():
                x, _ = eng.preprocess(images, self.backbone.size_divisibility)
                feats = self.backbone.forward_fmap(x)
                det = fcos.detect([feats[f] for f in fcos.in_features])
                probs, mask_scores = roi.run([feats[f] for f in roi.in_features], det, sizes)
                boxes, valid = eng.rescale_boxes(det["boxes"], sizes, out_sizes, det["count"])
                r_cap = det["boxes"].shape[1]
                R = n * r_cap
                rec = B("records_slot{}".format(slot), (n, r_cap, parallel.RECORD_FIELDS), torch.float32, False)
                parallel.pack_slots(rec, boxes, det["scores"], det["classes"], mask_scores, det["locations"], valid, det["count"])
                cand = B("records_cand{}".format(slot), tuple(det["cand_count"].shape), det["cand_count"].dtype, False)
                cand.copy_(det["cand_count"], non_blocking=True)
                masks = B("rec_masks{}".format(slot), (R, oh, ow), torch.uint8, False)
                lib.paste_masks(probs, boxes, valid, masks, R, probs.shape[-1], oh, ow, 0.5)
                moff = B("rle_mask_offset{}".format(slot), (R + 1,), torch.int64, False)
                runs = B("rle_runs{}".format(slot), (cap,), torch.int32, False)
                lib.rle_encode(masks, B("rle_col_count", (R, ow), torch.int32, False), B("rle_col_offset", (R, ow), torch.int32, False),
                               B("rle_total", (R,), torch.int32, False), moff, B("rle_positions", (cap,), torch.int32, False), runs,
                               boxes, valid)
                return det, rec, cand, masks, moff, runs

            eng.trim()
            det, rec, cand, masks, moff, runs = eng.graphed(("records", self.graph_tokens(), sig, (oh, ow), slot, cap), plan,
                                                            keep=self.packed_refs())
            done = torch.cuda.Event()
            done.record()
            slot_free[slot] = done
            r_cap = det["boxes"].shape[1]
            R = n * r_cap
            ctx = dict(n=n, r_cap=r_cap, slot=slot, out_size=(oh, ow), masks=masks, runs=runs, cap=cap, cand_cap=det["cand_cap"],
                       h_rec=eng.pinned("h_records{}".format(slot), (n, r_cap, parallel.RECORD_FIELDS), torch.float32),
                       h_cand=eng.pinned("h_rcand{}".format(slot), tuple(cand.shape), torch.int32),
                       h_off=eng.pinned("h_rle_off{}".format(slot), (R + 1,), torch.int64))
            back.wait_event(done)
            with torch.cuda.stream(back):
                ctx["h_rec"].copy_(rec, non_blocking=True)
                ctx["h_off"].copy_(moff, non_blocking=True)
                ctx["h_cand"].copy_(cand, non_blocking=True)
                ctx["ready"] = torch.cuda.Event()
                ctx["ready"].record(back)
            pending.append(ctx)
            return True

        if not roi.mask_on:
            raise NotImplementedError("inference_records needs MODEL.MASK_ON")
        state["staged"] = stage(next(it, None), 0)
        for _ in range(depth):
            if not launch_next():
                break
        while pending:
            launch_next()
            ctx = pending.popleft()
            ctx["ready"].synchronize()
            if bool((ctx["h_cand"] > ctx["cand_cap"]).any()):
                raise RuntimeError("FCOS candidate buffer overflow (> {} candidates above threshold in one level)".format(ctx["cand_cap"]))
            R = ctx["n"] * ctx["r_cap"]
            n_runs = int(ctx["h_off"][R])
            while n_runs > ctx["cap"]:
                # the run buffers were too small for this batch: grow them (for every later batch too) and encode again --
                # the pasted masks of the slot are still there
                state["cap"] = max(2 * state["cap"], n_runs + (n_runs >> 2))
                moff = eng.buffer("rle_mask_offset{}".format(ctx["slot"]), (R + 1,), torch.int64, False)
                pos = eng.buffer("rle_positions", (state["cap"],), torch.int32, False)
                runs = eng.buffer("rle_runs{}".format(ctx["slot"]), (state["cap"],), torch.int32, False)
                lib.rle_encode(ctx["masks"], eng.buffer("rle_col_count", (R, ctx["out_size"][1]), torch.int32, False),
                               eng.buffer("rle_col_offset", (R, ctx["out_size"][1]), torch.int32, False),
                               eng.buffer("rle_total", (R,), torch.int32, False), moff, pos, runs)       # (whole-mask scan)
                torch.cuda.current_stream().synchronize()
                ctx.update(runs=runs, cap=state["cap"])
                ctx["h_off"].copy_(moff)
                n_runs = int(ctx["h_off"][R])
            h_runs = eng.pinned("h_rle_runs{}".format(ctx["slot"]), (max(ctx["cap"], 1),), torch.int32)
            with torch.cuda.stream(back):
                h_runs[:n_runs].copy_(ctx["runs"][:n_runs], non_blocking=True)
            back.synchronize()
            yield parallel.BatchResult(ctx["h_rec"], ctx["h_off"], h_runs[:n_runs], ctx["out_size"])

    @torch.no_grad()
    def inference(self, batched_inputs, detected_instances=None, do_postprocess=True):
        """list[{"image": [3,H,W] BGR (float or uint8), "height", "width"}] -> list[{"instances": Instances}].

        The whole batch runs on fixed-size device buffers (no host sync) up to and including the mask
        paste-back; one D2H copy of the detection counts / box validity then cuts out the per-image
        ``Instances`` (slices, no per-field gathers unless a box became empty after clipping)."""
        eng = runtime.engine_for(self.cfg)
        n = len(batched_inputs)
        sizes = [(int(b["image"].shape[-2]), int(b["image"].shape[-1])) for b in batched_inputs]
        roi = self.roi_heads
        if detected_instances is not None:
            images = [b["image"].to(eng.device, non_blocking=True) for b in batched_inputs]
            x, _ = eng.preprocess(images, self.backbone.size_divisibility)
            feats = self.backbone.forward_fmap(x)
            results = [i.to(eng.device) for i in detected_instances]
            results = roi.forward_with_given_boxes({k: v.nchw() for k, v in feats.items()}, results)
            if not do_postprocess:
                return results
            return [{"instances": self.detector_postprocess(inst, b.get("height", sz[0]), b.get("width", sz[1]))}
                    for inst, b, sz in zip(results, batched_inputs, sizes)]
        return self._finish(self._launch(batched_inputs, do_postprocess, False, 0))

    def _launch(self, batched_inputs, do_postprocess, inputs_resident, slot):
        """Enqueue the whole device side of one batch (inputs -> graph replay -> result snapshots -> paste-back) without
        waiting for it.  ``slot`` selects the set of pinned host buffers, so that two batches can be in flight."""
        eng = runtime.engine_for(self.cfg)
        eng.trim()
        n = len(batched_inputs)
        sizes = [(int(b["image"].shape[-2]), int(b["image"].shape[-1])) for b in batched_inputs]
        roi = self.roi_heads
        out_sizes = [(int(b.get("height", sz[0])), int(b.get("width", sz[1]))) for b, sz in zip(batched_inputs, sizes)]
        # inputs land in engine-owned buffers (static addresses: the launch plan below is replayed as a CUDA graph)
        sig = tuple((tuple(b["image"].shape), b["image"].dtype) for b in batched_inputs)
        images, _ = eng.image_buffers("input_image", sig)
        if not inputs_resident:                                         # inference_stream has put them there already
            for dst, b in zip(images, batched_inputs):
                dst.copy_(b["image"], non_blocking=True)
        fcos = self.proposal_generator

        def plan():
            x, _ = eng.preprocess(images, self.backbone.size_divisibility)
            feats = self.backbone.forward_fmap(x)
            det = fcos.detect([feats[f] for f in fcos.in_features])
            probs = mask_scores = boxes = valid = kps = None
            if roi.mask_on:
                probs, mask_scores = roi.run([feats[f] for f in roi.in_features], det, sizes)
            if roi.keypoint_on:                                        # center_heads.py:441-442: after the masks
                kps = roi.run_keypoints([feats[f] for f in roi.kp_in_features], det, sizes)
            if do_postprocess:
                boxes, valid = eng.rescale_boxes(det["boxes"], sizes, out_sizes, det["count"])
            return det, probs, mask_scores, boxes, valid, kps

        # the key names the weights the capture bakes in (module identity + weight generation): a model whose weights were
        # reloaded, or a second model built from the same cfg, can never replay another capture
        det, probs, mask_scores, boxes, valid, kps = eng.graphed(
            ("inference", self.graph_tokens(), sig, tuple(out_sizes), bool(do_postprocess)), plan, keep=self.packed_refs())
        r_cap = det["boxes"].shape[1]
        # one snapshot of the small per-detection tensors (the engine reuses its buffers on the next call)
        scores, classes, locs = det["scores"].clone(), det["classes"].clone(), det["locations"].clone()
        mscores = mask_scores.reshape(n, r_cap).clone() if mask_scores is not None else None
        keypoints = None
        if kps is not None:
            keypoints = torch.cat([kps[..., :2], kps[..., 3:4]], dim=-1)   # keypoint_head.py:116 (x, y, score); a copy
            if do_postprocess:                                          # detector_postprocess [d2]: x *= scale_x, y *= scale_y
                pp = eng.pp_params(sizes, out_sizes)
                keypoints[..., 0] *= pp[:, 0].view(n, 1, 1)
                keypoints[..., 1] *= pp[:, 1].view(n, 1, 1)
        h_count = eng.pinned("h_count{}".format(slot), (n,), torch.int32)
        h_cand = eng.pinned("h_cand{}".format(slot), tuple(det["cand_count"].shape), torch.int32)
        h_count.copy_(det["count"], non_blocking=True)
        h_cand.copy_(det["cand_count"], non_blocking=True)
        if do_postprocess:
            boxes = boxes.clone()
            h_valid = eng.pinned("h_valid{}".format(slot), (n, r_cap), torch.uint8)
            h_valid.copy_(valid, non_blocking=True)
        else:
            boxes = det["boxes"].clone()
            pm = probs.clone() if probs is not None else None
        # the host needs only the counts / validity flags: wait for those, and let the paste-back kernel (the last ~3% of
        # the device work) run while the Instances are assembled -- everything returned is stream-ordered device memory
        ready = torch.cuda.Event()
        ready.record()
        masks = None
        if do_postprocess and probs is not None:
            masks = eng.paste_batch(probs, boxes, valid, out_sizes, dtype=torch.bool)
        done = torch.cuda.Event()
        done.record()
        return dict(done=done, eng=eng, n=n, sizes=sizes, out_sizes=out_sizes, do_postprocess=do_postprocess, r_cap=r_cap, scores=scores,
                    classes=classes, locs=locs, mscores=mscores, boxes=boxes, masks=masks, pm=None if do_postprocess else pm,
                    ready=ready, h_count=h_count, h_cand=h_cand, h_valid=h_valid if do_postprocess else None,
                    cand_cap=det["cand_cap"], have_probs=probs is not None, keypoints=keypoints)

    def graph_tokens(self):
        return tuple(m.graph_token() for m in (self.backbone, self.proposal_generator, self.roi_heads))

    def packed_refs(self):
        """The packed weight objects of the three plug-ins (packing them if needed): kept alive by a captured graph."""
        return tuple(m._pack()[1] for m in (self.backbone, self.proposal_generator, self.roi_heads))

    def _finish(self, ctx):
        """Wait for the small result-size tensors of a launched batch and cut the per-image ``Instances`` out."""
        (eng, n, sizes, out_sizes, do_postprocess, r_cap, scores, classes, locs, mscores, boxes, masks, pm, ready, h_count, h_cand,
         h_valid, cand_cap, have_probs) = [ctx[k] for k in (
             "eng", "n", "sizes", "out_sizes", "do_postprocess", "r_cap", "scores", "classes", "locs", "mscores", "boxes", "masks",
             "pm", "ready", "h_count", "h_cand", "h_valid", "cand_cap", "have_probs")]
        ready.synchronize()
        counts = h_count.tolist()
        if bool((h_cand > cand_cap).any()):
            raise RuntimeError("FCOS candidate buffer overflow (> {} candidates above threshold in one level)".format(cand_cap))
        total = sum(counts)
        out = []
        for i, k in enumerate(counts):
            sel = slice(0, k)
            if do_postprocess:
                v = h_valid[i, :k]
                if not bool(v.all()):
                    sel = torch.nonzero(v).squeeze(1).to(eng.device)
                inst = Instances(tuple(out_sizes[i]))
            else:
                inst = Instances(tuple(sizes[i]))
            inst.pred_boxes = Boxes(boxes[i, sel])
            inst.scores = scores[i, sel]
            inst.pred_classes = classes[i, sel]
            inst.locations = locs[i, sel]
            if have_probs:
                inst.pred_masks = masks[i][sel] if do_postprocess else pm[i * r_cap:(i + 1) * r_cap][sel]
                if mscores is not None and total > 0:                    # center_heads.py:511-517
                    inst.mask_scores = mscores[i, sel]
            if ctx["keypoints"] is not None:
                inst.pred_keypoints = ctx["keypoints"][i, sel]
            out.append({"instances": inst} if do_postprocess else inst)
        return out

    def preprocess_image(self, batched_inputs):
        """[d2] API parity: returns an ``ImageList`` whose tensor is the normalised padded batch [N,3,H,W]."""
        eng = runtime.engine_for(self.cfg)
        images = [b["image"].to(eng.device) for b in batched_inputs]
        x, sizes = eng.preprocess(images, self.backbone.size_divisibility, fused_stem=False)
        return ImageList(x.nchw(), sizes)

    def detector_postprocess(self, results, output_height, output_width, mask_threshold=0.5):
        """detector_postprocess + paste_masks_in_image [d2] (fork restatement: deploy_utils.py:129-158)."""
        eng = runtime.engine_for(self.cfg)
        out = Instances((output_height, output_width))
        r = len(results)
        boxes = results.pred_boxes.tensor.contiguous()
        if results.has("pred_masks") and r > 0:
            b2, valid, masks = eng.paste(results.pred_masks.contiguous(), boxes, output_height, output_width,
                                         results.image_size, mask_threshold)
        else:
            b2 = boxes.clone()
            b2[:, 0::2] = (b2[:, 0::2] * (output_width / results.image_size[1])).clamp(0, output_width)
            b2[:, 1::2] = (b2[:, 1::2] * (output_height / results.image_size[0])).clamp(0, output_height)
            valid = ((b2[:, 2] - b2[:, 0]) > 0) & ((b2[:, 3] - b2[:, 1]) > 0)
            masks = None
        keep = valid.bool()
        out.pred_boxes = Boxes(b2[keep])
        for k, v in results.get_fields().items():
            if k in ("pred_boxes", "pred_masks"):
                continue
            out.set(k, v[keep])
        if out.has("pred_keypoints"):                                   # [d2]: x *= scale_x, y *= scale_y
            kp = out.pred_keypoints.clone()
            kp[:, :, 0] *= output_width / results.image_size[1]
            kp[:, :, 1] *= output_height / results.image_size[0]
            out.pred_keypoints = kp
        if masks is not None:
            out.pred_masks = masks[keep].bool()
        elif results.has("pred_masks"):
            out.pred_masks = torch.zeros((0, output_height, output_width), dtype=torch.bool, device=boxes.device)
        return out

    # -- the fork's export-friendly variant ------------------------------------------------------
    @torch.no_grad()
    def forward_tensor(self, img, hw=None):
        """``modified_class.py:27-40``: already normalised + padded ``img[1,3,H,W]`` -> 6-tuple
        (locations, mask_scores, pred_boxes, pred_classes, pred_masks, scores) of batch element 0.  ``hw``: the
        ``image_sizes`` list of the fork's ``FakeImageList`` (modified_class.py:11-24); like there it defaults to
        1344 x 1344 per image whatever the tensor's extent (the fork feeds its fixed 1344 x 1344 export shape)."""
        features = self.backbone(img)
        sizes = _Sizes(hw if hw is not None else [(1344, 1344)] * img.shape[0])
        proposals, _ = self.proposal_generator(sizes, features, None)
        results, _ = self.roi_heads(sizes, features, proposals, None)
        r = results[0]
        ms = r.mask_scores if r.has("mask_scores") else r.scores.new_zeros((0,))
        pm = r.pred_masks if r.has("pred_masks") else r.scores.new_zeros((0, 1, 28, 28))
        return r.locations, ms, r.pred_boxes.tensor, r.pred_classes, pm, r.scores


def _register():
    m = getattr(META_ARCH_REGISTRY, "_obj_map", getattr(META_ARCH_REGISTRY, "_map", None))
    if m is not None and "GeneralizedRCNN" in m:
        m.pop("GeneralizedRCNN")            # same move as the reference: convert_model_into_onnx.py:30-31
    META_ARCH_REGISTRY.register(GeneralizedRCNN)


_register()

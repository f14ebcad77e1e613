"""Label wire format and the label-production loop around the assign kernel.

Format (one line of text per text line, consumed by the masked-pretraining dataset):
    "<line_id> <l0> <l1> ... <ln>\\n"        labels of the frames whose image_mask == 1, space separated
  scripts/produce_kmeans_labels.py:83-85   (written while iterating),
  scripts/common.py:51-54 save_labels      (written from a dict at the end),
  scripts/produce_vqvae_labels.py:25-44    (compute_labels: {line_id: [labels]}).

The production loop keeps the device busy: labels of batch i travel device->host through a ring of pinned buffers
on a copy stream while batch i+1 is being assigned; the host formats and writes batch i-1 meanwhile.
"""
import numpy as np
import torch

from .kmeans_labels import KMeansLabeller


def format_label_line(line_id, line_labels):
    return f"{line_id} {' '.join([str(label) for label in line_labels])}\n"


def save_labels(data, path):
    """scripts/common.py:51-54."""
    with open(path, "w") as f:
        for line_id, line_labels in data.items():
            f.write(format_label_line(line_id, line_labels))


def load_labels(path):
    """Inverse of save_labels: {line_id: [int labels]}."""
    out = {}
    with open(path) as f:
        for line in f:
            parts = line.split()
            if parts:
                out[parts[0]] = [int(p) for p in parts[1:]]
    return out


class LabelWriter:
    """Streaming writer of the label file (produce_kmeans_labels.py:83-85)."""

    def __init__(self, path):
        self.path = path
        self.file = open(path, "w")
        self.lines_written = 0

    def write_batch(self, ids, image_masks, assignment):
        """assignment [B, T] integer array (host); image_masks [B, T] {0,1}; ids: B line ids."""
        assignment = np.asarray(assignment)
        for line_id, line_image_mask, line_ids in zip(ids, image_masks, assignment):
            line_ids = line_ids[np.asarray(line_image_mask) == 1]
            self.file.write(format_label_line(line_id, line_ids))
            self.lines_written += 1

    def close(self):
        self.file.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _HostRing:
    """`depth` pinned int64 staging buffers + events for asynchronous label read-back."""

    def __init__(self, depth, device):
        self.depth, self.device = depth, device
        self.bufs = [None] * depth
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.stream = torch.cuda.Stream(device=device)

    def stage(self, slot, labels_dev):
        n = labels_dev.numel()
        if self.bufs[slot] is None or self.bufs[slot].numel() < n:
            self.bufs[slot] = torch.empty(max(n, 4096), dtype=torch.int64).pin_memory()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.bufs[slot][:n].copy_(labels_dev.reshape(-1), non_blocking=True)
            self.events[slot].record(self.stream)
        labels_dev.record_stream(self.stream)
        return n

    def fetch(self, slot, n, shape):
        self.events[slot].synchronize()
        return self.bufs[slot][:n].numpy().reshape(shape).copy()


def produce_kmeans_labels(batches, centers, output_path, depth=3):
    """The loop of scripts/produce_kmeans_labels.py:compute_features with the model call left to the caller:
    `batches` yields dicts with 'features' ([B, D, T] or [B, D, 1, T] CUDA tensor, what ``model(images)`` returns,
    :52-57), 'ids' and 'image_masks'.  `centers` [K, D] (CUDA tensor or numpy, what ``np.load(kmeans_path)`` gives).
    Returns the number of lines written."""
    first = None
    labeller = None
    pending = []            # (slot, n, shape, ids, image_masks)
    with LabelWriter(output_path) as writer:
        ring = None
        for i, batch in enumerate(batches):
            feats = batch["features"]
            if labeller is None:
                c = torch.as_tensor(centers, dtype=torch.float32).to(feats.device)
                labeller = KMeansLabeller(c)
                ring = _HostRing(depth, feats.device)
            assignment = labeller.assign_features(feats)                   # [B, T] int64 on the device
            slot = i % depth
            if len(pending) == depth:                                      # the slot is about to be reused: drain it
                s, n, shape, ids, masks = pending.pop(0)
                writer.write_batch(ids, masks, ring.fetch(s, n, shape))
            n = ring.stage(slot, assignment)
            pending.append((slot, n, tuple(assignment.shape), batch["ids"], batch["image_masks"]))
        for s, n, shape, ids, masks in pending:
            writer.write_batch(ids, masks, ring.fetch(s, n, shape))
        return writer.lines_written


def compute_labels(model, batches):
    """scripts/produce_vqvae_labels.py:25-44: {line_id: labels of the frames with image_mask == 1} from a VQVAE's
    quantizer.  `batches` yields dicts with 'images' (already on the model's device), 'ids', 'image_masks'."""
    data = {}
    with torch.no_grad():
        for batch in batches:
            feats = model.encode(batch["images"])
            if hasattr(model, "labels"):          # labels only: no quantized output, no decoder projection
                labels = model.labels(feats)
                N, T = feats.shape[0], feats.shape[3]              # (the encoders of the reference emit height 1)
                if labels.numel() != N * T:
                    T = labels.numel() // N
            else:
                tokens, labels = model.quantize(feats)
                N, _, _, T = tokens.shape
            labels = labels.reshape(N, T).cpu().numpy()
            for line_id, line_image_mask, line_labels in zip(batch["ids"], batch["image_masks"], labels):
                data[line_id] = line_labels[np.asarray(line_image_mask) == 1].tolist()
    return data

"""Synthetic JPEG files for the decoder tests (Pillow as the encoder): a smooth colour field plus noise, so that the
Huffman streams carry varied run / size symbols."""
import io

import numpy as np


def _jpeg_bytes(h, w, seed=0, gray=False, **kw):
    from PIL import Image
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 9.0 + yy / 13.0), 128 + 90 * np.cos(xx / 7.0), 128 + 80 * np.sin(yy / 5.0)], -1)
    img = np.clip(img + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img[..., 0] if gray else img).save(buf, format="JPEG", **kw)
    return buf.getvalue()


def _cmyk_jpeg():
    """A 4-component (CMYK) JPEG: valid, but outside the decoder's subset."""
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.full((16, 16, 4), 100, dtype=np.uint8), mode="CMYK").save(buf, format="JPEG")
    return buf.getvalue()

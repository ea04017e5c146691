"""Synthesise scenes/textures/earthmap.jpg (1024x512 equirectangular, RGB).

The reference ships a NASA-derived earth map under the same path; the texel VALUES are irrelevant to
the path being accelerated (ImageTexture::value is a nearest-texel fetch, src/world/texture.rs:98-117)
and no reference test pins them, so the fixture here is a procedural stand-in of the same size:
low-frequency "continents" over "ocean" plus polar caps, deterministic (no RNG).
"""
import os

import numpy as np
from PIL import Image

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes", "textures", "earthmap.jpg")


def main():
    w, h = 1024, 512
    lon = (np.arange(w) + 0.5) / w * 2 * np.pi
    lat = (0.5 - (np.arange(h) + 0.5) / h) * np.pi
    lon, lat = np.meshgrid(lon, lat)
    x, y, z = np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)
    f = np.zeros_like(x)
    # a fixed set of plane waves on the sphere: smooth, seamless at the date line
    dirs = [(1.0, 0.3, 0.2, 2.1), (-0.4, 1.0, 0.5, 3.3), (0.2, -0.7, 1.0, 4.7), (0.9, 0.8, -0.3, 6.1),
            (-0.6, 0.1, -1.0, 7.9), (0.3, 0.9, 0.9, 11.3), (-1.0, -0.5, 0.4, 13.7)]
    for k, (a, b, c, fr) in enumerate(dirs):
        f += np.sin(fr * (a * x + b * y + c * z) + 0.7 * k) / (1.0 + 0.35 * k)
    land = f > 0.25
    img = np.zeros((h, w, 3), np.float64)
    depth = np.clip((0.25 - f) / 2.0, 0, 1)
    img[..., 0] = 10 + 20 * (1 - depth)
    img[..., 1] = 40 + 60 * (1 - depth)
    img[..., 2] = 110 + 90 * (1 - depth)
    elev = np.clip((f - 0.25) / 1.5, 0, 1)
    dry = np.clip(1.0 - np.abs(lat) / 0.9, 0, 1)
    lr = 60 + 120 * elev + 70 * dry * (1 - elev)
    lg = 110 + 40 * elev - 10 * dry
    lb = 50 + 60 * elev
    for ch, v in enumerate((lr, lg, lb)):
        img[..., ch] = np.where(land, v, img[..., ch])
    ice = np.abs(lat) > (1.25 - 0.1 * np.sin(5 * lon))
    img[ice] = (235, 240, 245)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    Image.fromarray(np.clip(img, 0, 255).astype(np.uint8), "RGB").save(OUT, quality=90)


if __name__ == "__main__":
    main()

"""Asset conversion for the scene loader: VTK PolyData (.vtp) and STL meshes -> Wavefront OBJ,
and indexed ("welded") triangle meshes <-> the un-welded triangle soup the simulator consumes.

The reference prepares its IRCAD organs with utils/vtp_to_obj.py (vtk.vtkXMLPolyDataReader, prints the corner coordinates of
every polygon, one polygon per line); this module needs no vtk: it parses the .vtp XML itself (ascii, inline base64 -- raw or
zlib-compressed -- and appended raw / base64 data arrays; little-endian) and writes real OBJ files that
tinyobj / objloader.h (and mcrt_load_obj) read back triangle for triangle, polygons as fans.

    python -m mcray_tracing_b200.convert organ.vtp organ.obj [--dump]
    python -m mcray_tracing_b200.convert organ.stl organ.obj [--no-weld]
"""
from __future__ import annotations

import base64
import struct
import sys
import xml.etree.ElementTree as ET
import zlib
from pathlib import Path

import numpy as np

_VTK_DTYPES = {"Float32": "<f4", "Float64": "<f8", "Int8": "<i1", "UInt8": "<u1", "Int16": "<i2", "UInt16": "<u2", "Int32": "<i4",
               "UInt32": "<u4", "Int64": "<i8", "UInt64": "<u8"}


# ------------------------------------------------------------------------------------------------
# VTK XML PolyData
# ------------------------------------------------------------------------------------------------
def _decode_blocks(raw: bytes, header_dtype: str, compressed: bool, b64: bool) -> bytes:
    """one DataArray payload: [header][data]; header = byte count, or the zlib block table"""
    hsz = np.dtype(header_dtype).itemsize
    if b64:
        # the header and the data are base64-encoded separately when compressed, together otherwise; decode enough for the header first
        if compressed:
            first = base64.b64decode(raw[: ((3 * hsz + 2) // 3) * 4])
            nblocks = int(np.frombuffer(first[:hsz], header_dtype)[0])
            hbytes = (3 + nblocks) * hsz
            hlen = ((hbytes + 2) // 3) * 4
            head = np.frombuffer(base64.b64decode(raw[:hlen])[:hbytes], header_dtype)
            data = base64.b64decode(raw[hlen:])
            out, off = [], 0
            for csz in head[3:3 + nblocks]:
                out.append(zlib.decompress(data[off:off + int(csz)])); off += int(csz)
            return b"".join(out)
        blob = base64.b64decode(raw)
        n = int(np.frombuffer(blob[:hsz], header_dtype)[0])
        return blob[hsz:hsz + n]
    if compressed:
        nblocks = int(np.frombuffer(raw[:hsz], header_dtype)[0])
        head = np.frombuffer(raw[:(3 + nblocks) * hsz], header_dtype)
        off = (3 + nblocks) * hsz
        out = []
        for csz in head[3:3 + nblocks]:
            out.append(zlib.decompress(raw[off:off + int(csz)])); off += int(csz)
        return b"".join(out)
    n = int(np.frombuffer(raw[:hsz], header_dtype)[0])
    return raw[hsz:hsz + n]


def read_vtp(path) -> tuple[np.ndarray, list[np.ndarray]]:
    """-> (points [n, 3] float64, polygons: list of index arrays).  Triangle strips are expanded to triangles."""
    data = Path(path).read_bytes()
    appended_raw = None
    marker = data.find(b"<AppendedData")
    if marker >= 0:
        enc_raw = b'encoding="raw"' in data[marker:marker + 200]
        start = data.find(b"_", data.find(b">", marker)) + 1
        end = data.rfind(b"</AppendedData>")
        appended_raw = (data[start:end], enc_raw)
        # raw appended bytes are not valid XML: cut them out before parsing
        data = data[:start] + data[end:]
    root = ET.fromstring(data)
    if root.get("byte_order", "LittleEndian") != "LittleEndian":
        raise ValueError("read_vtp: only little-endian files are supported")
    header_dtype = "<u8" if root.get("header_type", "UInt32") == "UInt64" else "<u4"
    compressed = "ZLib" in (root.get("compressor") or "")
    if root.get("compressor") and not compressed:
        raise ValueError("read_vtp: unsupported compressor " + root.get("compressor"))

    def array(node) -> np.ndarray:
        dt = _VTK_DTYPES[node.get("type")]
        fmt = node.get("format", "ascii")
        if fmt == "ascii":
            return np.array((node.text or "").split(), dtype=np.float64 if "f" in dt else np.int64)
        if fmt == "binary":
            return np.frombuffer(_decode_blocks("".join((node.text or "").split()).encode(), header_dtype, compressed, True), dt)
        if fmt == "appended":
            if appended_raw is None:
                raise ValueError("read_vtp: appended DataArray without an AppendedData section")
            blob, is_raw = appended_raw
            off = int(node.get("offset"))
            if is_raw:
                return np.frombuffer(_decode_blocks(blob[off:], header_dtype, compressed, False), dt)
            return np.frombuffer(_decode_blocks(bytes(blob[off:]).strip(), header_dtype, compressed, True), dt)
        raise ValueError("read_vtp: unknown DataArray format " + fmt)

    pts_all, polys = [], []
    base = 0
    for piece in root.iter("Piece"):
        pnode = piece.find("Points/DataArray")
        pts = array(pnode).astype(np.float64).reshape(-1, 3)
        for tag, strip in (("Polys", False), ("Strips", True)):
            sec = piece.find(tag)
            if sec is None:
                continue
            arrs = {a.get("Name"): array(a).astype(np.int64) for a in sec.findall("DataArray")}
            conn, offs = arrs.get("connectivity"), arrs.get("offsets")
            if conn is None or offs is None or len(offs) == 0:
                continue
            start = 0
            for end in offs:
                cell = conn[start:end] + base
                start = int(end)
                if strip:
                    for k in range(len(cell) - 2):
                        polys.append(cell[[k, k + 1, k + 2]] if k % 2 == 0 else cell[[k + 1, k, k + 2]])
                elif len(cell) >= 3:
                    polys.append(cell)
        pts_all.append(pts)
        base += len(pts)
    if not pts_all:
        raise ValueError("read_vtp: no <Piece> with points")
    return np.concatenate(pts_all), polys


# ------------------------------------------------------------------------------------------------
# STL
# ------------------------------------------------------------------------------------------------
def read_stl(path) -> np.ndarray:
    """-> triangle soup [n, 3, 3] float32 (binary or ascii STL)"""
    data = Path(path).read_bytes()
    if len(data) >= 84:
        n = struct.unpack_from("<I", data, 80)[0]
        if 84 + 50 * n == len(data):                      # binary: the size field decides, not the "solid" prefix
            rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
            return rec["v"].astype(np.float32)
    toks = data.decode("ascii", errors="replace").split()
    v = [float(toks[i + 1 + k]) for i, t in enumerate(toks) if t == "vertex" for k in range(3)]
    if len(v) % 9:
        raise ValueError("read_stl: malformed ascii STL")
    return np.asarray(v, np.float32).reshape(-1, 3, 3)


# ------------------------------------------------------------------------------------------------
# indexed <-> soup
# ------------------------------------------------------------------------------------------------
def weld(soup: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """triangle soup [n, 3, 3] -> (vertices [m, 3], triangles [n, 3]); bit-identical corners share an index"""
    flat = np.ascontiguousarray(soup, np.float32).reshape(-1, 3)
    keys = flat.view(np.dtype((np.void, 12))).ravel()
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first)                             # vertices in order of first appearance
    rank = np.empty_like(order); rank[order] = np.arange(len(order))
    return flat[first[order]], rank[inv].reshape(-1, 3).astype(np.int64)


def indexed_to_soup(vertices: np.ndarray, polygons) -> np.ndarray:
    """(vertices [m, 3], polygons: [n, 3] array or list of index arrays) -> soup [t, 3, 3] float32; polygons become fans
    (v0, v(k-1), vk), the triangulation of tiny_obj_loader.cpp:272-285"""
    V = np.asarray(vertices, np.float64)
    tris = []
    for cell in polygons:
        cell = np.asarray(cell, np.int64)
        for k in range(2, len(cell)):
            tris.append((cell[0], cell[k - 1], cell[k]))
    T = np.asarray(tris, np.int64).reshape(-1, 3)
    return V[T].astype(np.float32)


def write_obj(path, vertices: np.ndarray, polygons, comment: str = "") -> None:
    """Wavefront OBJ with %.9g coordinates (float32 round-trips exactly) and 1-based faces"""
    with open(path, "w") as f:
        if comment:
            f.write("# " + comment + "\n")
        for v in np.asarray(vertices, np.float64):
            f.write("v %.9g %.9g %.9g\n" % (v[0], v[1], v[2]))
        for cell in polygons:
            f.write("f " + " ".join(str(int(i) + 1) for i in cell) + "\n")


def convert(src, dst, weld_vertices: bool = True) -> int:
    """.vtp / .stl -> .obj; returns the number of triangles the loader will see"""
    src, dst = Path(src), Path(dst)
    ext = src.suffix.lower()
    if ext == ".vtp":
        pts, polys = read_vtp(src)
        write_obj(dst, pts.astype(np.float32), polys, f"converted from {src.name}")
        return sum(len(c) - 2 for c in polys)
    if ext == ".stl":
        soup = read_stl(src)
        if weld_vertices:
            v, t = weld(soup)
        else:
            v, t = soup.reshape(-1, 3), np.arange(3 * len(soup)).reshape(-1, 3)
        write_obj(dst, v, t, f"converted from {src.name}")
        return len(soup)
    raise ValueError(f"convert: unsupported input '{ext}' (.vtp or .stl)")


def dump_polygons(src, out=sys.stdout) -> None:
    """the reference's utils/vtp_to_obj.py output: the corner coordinates of every polygon, one polygon per line"""
    pts, polys = read_vtp(src)
    for cell in polys:
        out.write(" ".join("%s %s %s" % (repr(float(pts[i][0])), repr(float(pts[i][1])), repr(float(pts[i][2]))) for i in cell) + " \n")


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if "--dump" in argv:
        argv.remove("--dump")
        dump_polygons(argv[0])
        return 0
    no_weld = "--no-weld" in argv
    if no_weld:
        argv.remove("--no-weld")
    if len(argv) != 2:
        print("usage: python -m mcray_tracing_b200.convert <in.vtp|in.stl> <out.obj> [--no-weld] | <in.vtp> --dump", file=sys.stderr)
        return 2
    n = convert(argv[0], argv[1], weld_vertices=not no_weld)
    print(f"{argv[1]}: {n} triangles")
    return 0


if __name__ == "__main__":
    sys.exit(main())

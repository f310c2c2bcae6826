#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- record how libhdf5 laid out the reference's shipped EMD file
(ExampleSpecimens/Au_cubeoctahedron_emd/Auparticle.emd) as a small JSON fixture, so that the EMD
writer test can compare structures where /root/reference is not mounted:

    python tools/make_emd_fixture.py      ->  tests/golden/emd_structure.json

Recorded: every path with kind / dtype / attribute names in creation order, the superblock fields,
the parameters the file carries (for the reader test), and the raw bytes of the dataspace /
datatype / fill-value messages of one float32 and one int32 dataset and of one string, one float32,
one int32 and one uint8 attribute message."""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import h5min  # noqa: E402

SRC = pathlib.Path("/root/reference/ExampleSpecimens/Au_cubeoctahedron_emd/Auparticle.emd")


def message_bytes(f, node, mtype, index=0):
    hits = [(p, s) for t, p, s in node.messages if t == mtype]
    p, s = hits[index]
    return f.b[p - 8:p + s].hex()


def attr_bytes(f, node, name):
    for t, p, s in node.messages:
        if t == 0x0C and f._attribute(p)[0] == name:
            return f.b[p - 8:p + s].hex()
    raise KeyError(name)


def main():
    f = h5min.File(SRC)
    out = {"source": str(SRC.relative_to("/root/reference")),
           "superblock": {"version": f.sb_version, "leaf_k": f.leaf_k, "internal_k": f.internal_k,
                          "size_offsets": f.size_offsets, "size_lengths": f.size_lengths, "base": f.base},
           "paths": {}}
    for path, node in f.root.walk():
        kind = "dataset" if node.data is not None else "group"
        out["paths"][path] = {"kind": kind, "dtype": str(node.data.dtype) if node.data is not None else None,
                              "rank": node.data.ndim if node.data is not None else None,
                              "attrs": list(node.attrs),
                              "attr_types": {k: (type(v).__name__ if not hasattr(v, "dtype") else str(v.dtype) + str(list(v.shape)))
                                             for k, v in node.attrs.items()}}
    zs = f.root["sample/atomic_numbers"]
    xs = f.root["sample/x_coordinates"]
    out["messages"] = {
        "int32_dataset": {"dataspace_309": message_bytes(f, zs, 1), "datatype": message_bytes(f, zs, 3), "fill": message_bytes(f, zs, 5)},
        "float32_dataset": {"datatype": message_bytes(f, xs, 3), "fill": message_bytes(f, xs, 5),
                            "attr_units": attr_bytes(f, xs, "units")},
        "attr_float32_voltage_50000": attr_bytes(f, f.root["microscope"], "voltage"),
        "attr_int32_sample_size_x_320": attr_bytes(f, f.root["imaging"], "sample_size_x"),
        "attr_uint8_emd_group_type_1": attr_bytes(f, f.root["data/images"], "emd_group_type"),
        "attr_string_voltage_units": attr_bytes(f, f.root["microscope"], "voltage_units"),
    }
    dst = ROOT / "tests" / "golden" / "emd_structure.json"
    dst.write_text(json.dumps(out, indent=1))
    print("wrote", dst, len(out["paths"]), "paths")


if __name__ == "__main__":
    main()

"""Recipe: compile the reference's hot-path modules, from the sources where they lie under /root/reference, into
`oracle/_ref/reference_pyc.zip` (byte-code only; `oracle/_ref/` is git-ignored, not gpurun-ignored, so the archive
travels to the GPU box like the built libmzb200.so).

Test/bench infrastructure: `bench.py --impl reference` and the bench's `cpu_baseline` leg import the UNMODIFIED reference
classes from this archive (zipimport) to time `SelfPlay.play_game` on the host cores.  No reference source is copied into
the repository; the archive is the Python analogue of a reference compiled into `oracle/_ref/*.so`.

    python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
import importlib.util
import marshal
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "reference_pyc.zip")
REFERENCE_ROOT = os.environ.get("MZB_REFERENCE_ROOT", "/root/reference")
# the path SURVEY.md §8(a) names + what those modules import
MODULES = ["self_play.py", "models.py", "replay_buffer.py", "trainer.py", "shared_storage.py",
           "games/abstract_game.py", "games/cartpole.py", "games/tictactoe.py", "games/connect4.py", "games/gomoku.py"]


def build(out=OUT, root=REFERENCE_ROOT):
    if not os.path.isfile(os.path.join(root, "self_play.py")):
        return None
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tmp = out + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("games/", b"")        # directory entry: `games` is a namespace package in the reference (no __init__.py)
        for rel in MODULES:
            src = os.path.join(root, rel)
            with open(src, "rb") as f:
                code = compile(f.read(), "reference/" + rel, "exec", dont_inherit=True, optimize=0)
            st = os.stat(src)
            # legacy (sourceless) layout: <module>.pyc next to where the .py would be; timestamp-based header
            data = (importlib.util.MAGIC_NUMBER + (0).to_bytes(4, "little") + int(st.st_mtime).to_bytes(4, "little")
                    + (st.st_size & 0xFFFFFFFF).to_bytes(4, "little") + marshal.dumps(code))
            z.writestr(rel[:-3] + ".pyc", data)
        z.writestr("PYTHON_VERSION", sys.version)
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build() or f"reference not present at {REFERENCE_ROOT}: nothing built")

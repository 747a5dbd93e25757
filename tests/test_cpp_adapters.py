"""The C++ drop-in adapters (include/orb_slam2/*.h: ORBextractor, ORBmatcher, CeresOptimizer with the reference's
names and call shapes) compile against libcmos_b200.so; without a GPU they fail loudly, with one they run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ceres_mono_orb_slam2_b200")


def _build(tmp):
    exe = os.path.join(tmp, "adapter_smoke")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cpp", "adapter_smoke.cpp"),
                    "-o", exe, "-L" + LIBDIR, "-lcmos_b200", "-Wl,-rpath," + LIBDIR], check=True, capture_output=True)
    return exe


def _have_gpu():
    import torch
    return torch.cuda.is_available()


def test_adapters_compile_and_refuse_to_run_without_gpu(tmp_path):
    exe = _build(str(tmp_path))
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 10 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_adapters_run_on_gpu(tmp_path):
    exe = _build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ADAPTERS_OK" in r.stdout, r.stdout + r.stderr


def test_essential_graph_edge_collection_follows_the_reference_order(tmp_path):
    """CeresOptimizer::CollectEssentialGraphEdges (host logic of the adapter) against the rules of CeresOptimizer.cc:793-895
    worked out by hand for a map that exercises each of them: loop connections first (weight >= 100 or the current / loop
    pair), then per keyframe in map order its parent, its OLDER loop edges, and its co-visible keyframes that are older and
    neither parent, child, loop edge nor already inserted."""
    exe = os.path.join(str(tmp_path), "essential_edges")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cpp", "essential_edges.cpp"),
                    "-o", exe, "-L" + LIBDIR, "-lcmos_b200", "-Wl,-rpath," + LIBDIR], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = [tuple(int(v) for v in line.split()) for line in out.strip().splitlines()]
    expected = [
        # loop connections, keyframes in index order (6 before 7)
        (1, 6, 0),
        (1, 7, 0), (0, 7, 0),
        # per keyframe: parent, older loop edges, older co-visible keyframes
        (0, 1, 1),                                   # keyframe 1: parent 0
        (3, 2, 1), (3, 2, 1),                        # keyframe 2 (id 30): parent 3, loop edge to 3 (id 20, older)
        (1, 3, 1),                                   # keyframe 3 (id 20): parent 1; loop edge 2 and co-visible 2 are younger
        (2, 4, 1), (3, 4, 1), (1, 4, 1),             # keyframe 4: parent 2, co-visible 3 and 1 (2 parent, 5 child skipped)
        (4, 5, 1), (0, 5, 1), (3, 5, 1),             # keyframe 5: parent 4, loop edge 0, co-visible 3 (0 loop edge, 4 parent)
        (5, 6, 1),                                   # keyframe 6: parent 5; (1,6) already inserted, 7 is a child
        (6, 7, 1), (5, 7, 1),                        # keyframe 7: parent 6; (0,7) already inserted; co-visible 5
    ]
    assert got == expected, got


def test_ba_graph_collection_follows_the_reference(tmp_path):
    """CeresOptimizer::CollectLocalGraph / CollectGlobalGraph / CollectLocalResult (host logic of the adapter) on a map worked
    out by hand against CeresOptimizer.cc:349-406, 421-502, 567-598 (local) and :88-175 (global): the current keyframe and
    its non-bad co-visible keyframes are local (a bad neighbour is marked local and can never become fixed); keyframe id 0
    is local but constant; observers of local points that are not local become fixed; bad points / bad keyframes are
    skipped; a point whose observers are all skipped is dropped."""
    exe = os.path.join(str(tmp_path), "ba_graph_collect")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cpp", "ba_graph_collect.cpp"),
                    "-o", exe, "-L" + LIBDIR, "-lcmos_b200", "-Wl,-rpath," + LIBDIR], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    got = {" ".join(l.split()[:2]): l.split()[2:] for l in out}
    # local: keyframe 2 (current), 1, 0 (id 0 -> constant, flag 1); bad 3 skipped; fixed (flag 3): 4 (reached through local
    # point 2) then 5 (through point 1), the order in which the local points' observations reach them
    assert got["local keyframes"] == ["2:0", "1:0", "0:1", "4:3", "5:3"]
    # local points in first-seen order over the local keyframes' GetMapPointMatches(): kf2 -> 2, 0, (4 bad), 3; kf1 -> 1; kf0 -> -
    assert got["local points"] == ["2", "0", "3", "1"]
    assert got["local obs"] == [
        "(1,2,101,10,0.5)", "(2,2,200,0,1)", "(4,2,400,0,1)",         # point 2: keyframes 1, 2, 4
        "(0,0,0,0,1)", "(2,0,201,10,0.5)",                            # point 0: keyframes 0, 2
        "(2,3,203,30,0.5)", "(4,3,401,10,0.5)",                       # point 3: keyframe 3 is bad
        "(0,1,1,10,0.5)", "(1,1,100,0,1)", "(5,1,500,0,1)"]           # point 1: keyframes 0, 1, 5
    assert got["result erase"] == ["(2,2)", "(5,1)"]
    assert got["result keyframes"] == ["2", "1", "0"] and got["result points"] == ["2", "0", "3", "1"]
    # global: all non-bad keyframes in map order, id 0 constant; point 4 bad; point 5 keeps its observation by keyframe 6
    assert got["global keyframes"] == ["0:1", "1:0", "2:0", "4:0", "5:0", "6:0"]
    assert got["global points"] == ["0", "1", "2", "3", "5", "6"]
    assert len(got["global obs"]) == 2 + 3 + 3 + 2 + 1 + 1 and got["global obs"][-2:] == ["(6,5,600,0,1)", "(5,6,501,10,0.5)"]

"""The C++ host layer (Scene / Scenes.xml / cSPH::Reset / Drop / scene switching) against vectors
produced by the reference's own scene code (tests/golden/*_scenes.npz).  CPU only: the cSPH object
is built with device=-1, which never touches a GPU."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import REFERENCE_XML, ROOT, sha
from pibiti_b200 import host
from pibiti_b200.lib import SIMPARAMS_DTYPE

SKIP_FIELDS = ("ff2",)      # never initialised by the reference (Scene.cpp), zero here


def check_scene_table(s: host.CSph, g):
    assert s.num_scenes == len(g["params"])
    assert s.curScene == int(g["cur_scene"])
    for i in range(s.num_scenes):
        par = s.scene_params(i)
        for name in SIMPARAMS_DTYPE.names:
            if name not in SKIP_FIELDS:
                assert par[name].tobytes() == g["params"][i:i + 1][name].tobytes(), (i, s.scene_title(i), name)
        assert s.scene_extra(i).tobytes() == g["extra"][i].tobytes(), (i, "scene extras")


def check_reset_and_drop(s: host.CSph, g, indices):
    for i in indices:
        if str(g["reset_sha"][i]) == "skipped":
            continue
        n = s.select_scene(i)                                   # srand(1); UpdScene -> InitScene -> Reset
        assert n == int(g["num_particles"][i])
        live = s.params
        for name in SIMPARAMS_DTYPE.names:
            if name not in SKIP_FIELDS:
                assert live[name].tobytes() == g["live"][i:i + 1][name].tobytes(), (i, name)
        pos, vel = s.host_arrays()
        assert np.array_equal(pos[:8], g["reset_head"][i])
        assert sha(pos) + sha(vel) == str(g["reset_sha"][i]), (i, s.scene_title(i), "Reset")
        s.srand(7)
        s.Drop(False)
        e = s.Drop(True)
        pos, vel = s.host_arrays()
        assert sha(pos[:, :3]) + sha(vel) + f"{e:08x}" == str(g["drop_sha"][i]), (i, s.scene_title(i), "Drop")


def test_repo_scenes_match_reference_loader(golden_repo_scenes):
    s = host.CSph(device=-1)
    check_scene_table(s, golden_repo_scenes)
    small = [i for i in range(s.num_scenes) if int(golden_repo_scenes["num_particles"][i]) <= 1_100_000]
    check_reset_and_drop(s, golden_repo_scenes, small)


def test_repo_8m_tank_reset(golden_repo_scenes):
    s = host.CSph(device=-1)
    check_reset_and_drop(s, golden_repo_scenes, [s.scene_index("tank 8M drop")])


@pytest.mark.skipif(not REFERENCE_XML.exists(), reason="needs /root/reference/Scenes.xml")
def test_reference_scenes_xml_all_scenes(golden_ref_scenes):
    s = host.CSph(REFERENCE_XML, device=-1)
    assert s.num_scenes == 119
    check_scene_table(s, golden_ref_scenes)
    check_reset_and_drop(s, golden_ref_scenes, range(s.num_scenes))


def test_default_scene_constants():
    """SURVEY.md section 8 quotes these for scene 0."""
    s = host.CSph(device=-1)
    p = s.params
    assert int(p["numParticles"][0]) == 57344 and tuple(p["gridSize"][0]) == (50, 63, 50)
    assert int(p["numCells"][0]) == 157500 and int(p["maxParInCell"][0]) == 16
    assert np.float32(p["cellSize"][0][0]) == np.float32(0.008) and np.float32(p["h"][0]) == np.float32(0.01)
    assert np.float32(p["timeStep"][0]) == np.float32(0.0026)
    assert tuple(p["dyePos"][0]) == (0.0, 0.0, 0.0)             # InitScene overwrites it with App::dyePos
    assert tuple(s.scene_params(0)["dyePos"][0]) != (0.0, 0.0, 0.0)


def test_scene_navigation_and_chapters():
    s = host.CSph(device=-1)
    n = s.num_scenes
    chapters = [i for i in range(n) if s.scene_extra(i)[19] == 1.0]
    assert chapters[0] == 0                                     # scenes[0].bChapter forced true
    s.NextScene()
    assert s.curScene == 1
    s.PrevScene()
    s.PrevScene()
    assert s.curScene == n - 1                                  # wraps
    s.NextScene(chapter=True)
    assert s.curScene == chapters[0]
    s.NextScene(chapter=True)
    assert s.curScene == chapters[1]
    assert s.n == int(s.scene_params(s.curScene)["numParticles"][0])


def test_default_attribute_selects_start_scene(tmp_path):
    xml = tmp_path / "Scenes.xml"
    xml.write_text('<SPH><Scene name="a" ParticlesK="1"/><!-- <Scene name="x" def=""/> -->'
                   '<Scene name="b" ParticlesK="2" def=""/><Scene name="c" ParticlesK="1"/></SPH>')
    s = host.CSph(xml, device=-1)
    assert s.num_scenes == 3 and s.curScene == 1 and s.n == 2048


def test_missing_xml_gives_default_scene(tmp_path):
    s = host.CSph(tmp_path / "nope.xml", device=-1)
    assert s.num_scenes == 1 and s.n == 57344
    assert "cannot" in s.last_error().lower()


def test_xml_reader_edge_cases(tmp_path):
    xml = tmp_path / "Scenes.xml"
    xml.write_text("""<?xml version="1.0"?>
<!DOCTYPE SPH>
<SPH>
  <Options Windowed="0" WSizeX="800" WSizeY="600" VSyncOff="1" timAvgCnt="3" barsScale="10" showInfo="0"/>
  <!-- <Scene name="commented out" ParticlesK="99"/> -->
  <Scene name='single &amp; quoted'
         ParticlesK = "2"   World="0.3 0.3 0.3" unknownAttribute="ignored">
     text is ignored
     <Emitter size="3" pos="0.1 0.2 0.3" rot="10 20" vel="2.5"/>
     <Emitter size="4" size2="2"/>
     <Accel type="1" pos="0 0.1 0" size="0.1 0.1 0.1" acc="0 5 0"/>
  </Scene>
  <Scene name="hex" Particles="0x800"></Scene>
</SPH>""")
    s = host.CSph(xml, device=-1)
    assert s.num_scenes == 2
    assert s.scene_title(0) == "single & quoted"
    assert int(s.scene_params(0)["numParticles"][0]) == 2048
    ex = s.scene_extra(0)
    assert np.allclose(ex[20:28], [0.1, 0.2, 0.3, 10, 20, 2.5, 3, 0])
    assert ex[28 + 6] == 4 and ex[28 + 7] == 2
    acc = s.scene_params(0)["acc"][0]
    assert int(acc["type"][0]) == 1 and np.allclose(acc["acc"][0], [0, 5, 0]) and int(acc["type"][1]) == 0
    assert int(s.scene_params(1)["numParticles"][0]) == 0x800           # strtol base 0, like the reference's toInt
    assert host.load_options(xml) == {"Windowed": 0, "WSizeX": 800, "WSizeY": 600, "VSyncOff": 1, "timAvgCnt": 3,
                                      "barsScale": 10, "showInfo": 0}


def test_add_scene_xml_goes_through_the_scene_path():
    s = host.CSph(device=-1)
    idx = s.add_scene_xml('<Scene name="synthetic" ParticlesK="4" World="0.4 0.4 0.4" particleH="0.012" CellSize="0.01"/>')
    p = s.scene_params(idx)
    assert tuple(p["gridSize"][0]) == (40, 40, 40) and int(p["numParticles"][0]) == 4096
    assert np.float32(p["h2"][0]) == np.float32(0.012) * np.float32(0.012)
    s.select_scene(idx)
    assert s.n == 4096


def test_update_emitter_prologue():
    """Wave phase advance, dyeClear countdown and the emitter ring index (App/Update.cpp:9-97)."""
    s = host.CSph(device=-1)
    s.select_scene("mini waves")
    p0 = s.params
    assert np.float32(p0["rAngle"][0]) == np.float32(-3.141592654 / 2) and int(p0["dyeClear"][0]) == 2
    s.UpdateEmitter()
    p1 = s.params
    assert np.float32(p1["rAngle"][0]) == np.float32(p0["rAngle"][0]) + np.float32(4.5) * np.float32(p0["timeStep"][0])
    assert int(p1["dyeClear"][0]) == 1

    s.select_scene("mini emitter rain")
    n = s.n
    pos0, _ = s.host_arrays()
    s.UpdateEmitter()
    assert s.emitId == 9                                        # 3x3 emitter batch
    pos1, vel1 = s.host_arrays()
    assert not np.array_equal(pos0[:9], pos1[:9]) and np.array_equal(pos0[9:], pos1[9:])
    speed = np.linalg.norm(vel1[:9, :3], axis=1)
    assert np.allclose(speed, 2.0, rtol=1e-6)                   # EmitVel, rotated
    for _ in range(4):
        s.UpdateEmitter()                                       # 5th call: rain counter fires a random Drop
    assert s.emitId > 45 and s.emitId < n


def test_checkpoint_roundtrip_host_only(tmp_path):
    s = host.CSph(device=-1)
    s.select_scene("mini emitter rain")
    for _ in range(3):
        s.UpdateEmitter()
    pos, vel = s.host_arrays()
    par, eid = s.params, s.emitId
    s.SaveState(tmp_path / "state.bin")
    t = host.CSph(device=-1)
    t.LoadState(tmp_path / "state.bin")
    assert t.n == s.n and t.emitId == eid and t.params.tobytes() == par.tobytes()
    p2, v2 = t.host_arrays()
    assert np.array_equal(p2, pos) and np.array_equal(v2, vel)
    with pytest.raises(Exception):
        t.LoadState(tmp_path / "missing.bin")


# ---- the per-step host prologue against the reference's own App::UpdateEmitter -------------
def _ref_host_lib():
    import ctypes as C
    from oracle import oracle as orc
    if not orc.available("reference"):
        pytest.skip("oracle/_ref/libsphref.so not built (needs the reference tree at build time)")
    L = C.CDLL(str(orc.REF_LIB))
    if not hasattr(L, "refh_update_emitter"):
        pytest.skip("libsphref.so predates refh_update_emitter: rebuild where the reference tree exists")
    L.refh_set_emitter.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
    L.refh_set_targets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return L


PROLOGUE_SCENES = ["mini emitter rain", "mini rotor Y", "mini propeller pair", "mini waves", "mini collider accel"]


@pytest.mark.parametrize("title", PROLOGUE_SCENES)
def test_update_emitter_matches_the_reference_text(title):
    """cSPH::UpdateEmitter against App::UpdateEmitter (source/App/Update.cpp:9-97), compiled from the reference
    tree into oracle/_ref/libsphref.so: 200 steps, bit-for-bit -- SimParams (rotor / wave angles, collider,
    accelerator and dye lag, dyeClear), the emitter ring index, the rain counter, the dirty flag and every particle
    slot the emitters and the rain drops write.  Both sides use the C library's rand(): they run one after the other
    from the same seed."""
    import ctypes as C
    from pibiti_b200.lib import SIMPARAMS_DTYPE
    steps = 200
    xml_dir = host.DEFAULT_SCENES_XML.parent

    def vp(a):
        return a.ctypes.data_as(C.c_void_p)

    def drive(step, setters):
        """the UI edits both sides receive: collider / dye / accelerator targets move, a second emitter switches on"""
        set_targets, set_emitter = setters
        if step == 10:
            set_targets(np.array([0.03, -0.05, 0.02, 0], np.float32), np.array([0.01, -0.06, 0.0], np.float32),
                        np.array([-0.02, -0.07, 0.03], np.float32))
        if step == 60:
            set_emitter(1, np.array([-0.04, 0.06, 0.02], np.float32), np.array([-35.0, 110.0], np.float32), 1.5, 4, 2)
        if step == 120:
            set_targets(np.array([-0.05, -0.09, -0.04, 0], np.float32), None, None)

    def clean(block):
        out = np.zeros(1, SIMPARAMS_DTYPE)
        for name in SIMPARAMS_DTYPE.names:
            out[name] = block[name]
        out["ff2"] = 0
        return out.tobytes()

    # -- this repository's host layer
    s = host.CSph(device=-1)
    s.select_scene(title)                   # srand(1) + InitScene
    ours = []
    for k in range(steps):
        drive(k, (s.set_targets, s.set_emitter))
        s.UpdateEmitter()
        pos, vel = s.host_arrays()
        ours.append((clean(s.params), s.emitId, s.cntRain, s.changed_flag(clear=(k % 7 == 0)), sha(pos[:, :3]), sha(vel[:, :3])))
    s.close()

    # -- the reference's text
    L = _ref_host_lib()
    import re
    text = re.sub(r"<!--.*?-->", "", (xml_dir / "Scenes.xml").read_text(errors="replace"), flags=re.S)
    titles = re.findall(r"<Scene\s+name=\"([^\"]*)\"", text)      # the reference does not read `name` on Linux (Scene_Load.cpp:37-39)
    L.refh_load(str(xml_dir).encode())
    L.refh_select_scene(titles.index(title))                       # srand(1) + InitScene
    L.refh_changed_flag(1)
    n = s_n = None
    for k in range(steps):
        drive(k, (lambda c, d, a: L.refh_set_targets(vp(c), None if d is None else vp(d), None if a is None else vp(a)),
                  lambda e, p, r, v, sz, sz2: L.refh_set_emitter(e, vp(p), vp(r), v, sz, sz2)))
        L.refh_update_emitter()
        par = np.zeros(1, SIMPARAMS_DTYPE)
        L.refh_live_params(vp(par))
        n = int(par["numParticles"][0])
        pos, vel = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
        L.refh_get_host(vp(pos), vp(vel))
        ref = (clean(par), int(L.refh_emit_id()), int(L.refh_cnt_rain()), bool(L.refh_changed_flag(int(k % 7 == 0))),
               sha(pos[:, :3]), sha(vel[:, :3]))     # w: the reference leaves vel4.w uninitialised (Update.cpp:76)
        for name, a, b in zip(("SimParams", "emitId", "cntRain", "bChangedAny", "positions", "velocities"), ours[k], ref):
            assert a == b, f"{title}: {name} differs from the reference at step {k}"
    if title == "mini emitter rain":
        assert L.refh_set_array_calls() > 2 * steps and ours[-1][2] != ours[3][2] or ours[-1][1] != 9   # emitters and rain ran

//! What `src/rendering.rs:21-69` (`pub fn render_scene(scene: &Scene) -> Vec<u8>`) becomes once the pixel loop runs in
//! librt_b200.  UNCOMPILED in this repository's image (no cargo/rustc): see rust/README.md.  It is written against the
//! reference's own types (`crate::scene::{Scene, Primitive}`, `crate::geometry::Shape3D`) and the `rt-sys` crate next to it;
//! the call site `src/main.rs:55` does not change.
use crate::geometry::Shape3D;
use crate::scene::{Primitive, Scene};
use rt_sys::*;

fn check(rc: std::os::raw::c_int) {
    if rc != RT_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(rt_last_error()) }.to_string_lossy().into_owned();
        panic!("librt_b200: error {rc}: {msg}"); // the reference's failure mode is a panic (unwrap / assert!), kept at the seam
    }
}

struct Flat {
    kind: Vec<i32>,
    shape: Vec<f64>,
    nrm: Vec<f64>,
    mat: Vec<f64>,
    emi: Vec<f64>,
    pos: Vec<f64>,
    rot: Vec<f64>,
    ior: Vec<f64>,
    mk: Vec<i32>,
}

/// Flattens `Primitive`s (scene.rs:13-20) into the arrays of `RtSceneDesc2`.  Finite primitives come from
/// `scene.bvh_finite_primitives.primitives` (any order: the library builds its own BVH).  `scene.infinite_primitives`
/// (scene.rs:37) is empty at HEAD and HEAD's `Shape3D` cannot express a plane, so nothing is pushed for it.
fn flatten(prims: &[Primitive]) -> Flat {
    let mut f = Flat { kind: vec![], shape: vec![], nrm: vec![], mat: vec![], emi: vec![], pos: vec![], rot: vec![], ior: vec![], mk: vec![] };
    for p in prims {
        match &p.object3d.shape {
            Shape3D::Triangle { a, b, c, a_norm, b_norm, c_norm } => {
                f.kind.push(RT_SHAPE_TRIANGLE);
                for x in [a, b, c] {
                    f.shape.extend_from_slice(x.as_slice());
                }
                for x in [a_norm, b_norm, c_norm] {
                    f.nrm.extend_from_slice(x.as_slice());
                }
            }
            Shape3D::Box { s } => {
                f.kind.push(RT_SHAPE_BOX);
                f.shape.extend_from_slice(s.as_slice());
                f.shape.extend_from_slice(&[0.0; 6]);
                f.nrm.extend_from_slice(&[0.0; 9]);
            }
        }
        f.mat.extend_from_slice(p.material.base_color_factor.as_slice());
        f.mat.push(p.material.metallic_factor);
        f.mat.push(p.material.metallic_roughness);
        f.emi.extend_from_slice(p.emission.as_slice());
        f.pos.extend_from_slice(p.object3d.position.as_slice());
        let q = p.object3d.rotation.quaternion(); // nalgebra stores (i, j, k, w)
        f.rot.extend_from_slice(&[q.i, q.j, q.k, q.w]);
        f.ior.push(p.ior);
        f.mk.push(RT_MATERIAL_PBR);
    }
    f
}

pub fn render_scene(scene: &Scene) -> Vec<u8> {
    let f = flatten(&scene.bvh_finite_primitives.primitives);
    let desc = RtSceneDesc2 {
        base: RtSceneDesc {
            width: scene.width,
            height: scene.height,
            samples: scene.samples,
            ray_depth: scene.ray_depth,
            bg_color: scene.bg_color.into(),
            camera_position: scene.camera_position.into(),
            camera_forward: scene.camera_forward.into(),
            camera_right: scene.camera_right.into(),
            camera_up: scene.camera_up.into(),
            camera_fov_x: scene.camera_fov_x,
            camera_fov_y: scene.camera_fov_y,
            n_tris: f.kind.len() as i32,
            reserved0: 0,
            tri_v: f.shape.as_ptr(),
            tri_n: f.nrm.as_ptr(),
            tri_material: f.mat.as_ptr(),
            tri_emission: f.emi.as_ptr(),
        },
        shape_kind: f.kind.as_ptr(),
        position: f.pos.as_ptr(),
        rotation: f.rot.as_ptr(),
        ior: f.ior.as_ptr(),
        material_kind: f.mk.as_ptr(),
    };
    let mut handle = std::ptr::null_mut();
    let mut out = vec![0u8; (scene.width * scene.height * 3) as usize];
    unsafe {
        check(rt_scene_create2(&desc, 0, &mut handle));
        check(rt_render(handle, &RtRenderParams::default(), out.as_mut_ptr(), std::ptr::null_mut()));
        rt_scene_destroy(handle);
    }
    out // W*H*3 RGB8, row 0 = top: exactly what dump_rendered_to_ppm (main.rs:88-95) writes
}

//! The engine behind bevy_ggrs' own names.  SOURCE ONLY — never compiled in the build image (no Rust toolchain).
//!
//! What a user changes in `examples/stress_tests/particles.rs`:
//!
//! ```ignore
//! app.add_plugins(GgrsPlugin::<Config>::default())
//!    .add_plugins(B200RollbackPlugin { max_entities: 1_000_000, max_depth: 9 })   // new
//!    .rollback_component_with_clone_b200::<TransformPod>()                         // was rollback_component_with_clone::<Transform>()
//!    .rollback_component_with_copy_b200::<Velocity>()
//!    .rollback_component_with_copy_b200::<Ttl>()
//!    .checksum_component_b200::<Velocity>(0, 12, true)                             // was checksum_component_with_hash::<Velocity>()
//!    .checksum_component_b200::<TransformPod>(0, 12, true)                         // was the translation-bits closure
//!    .add_gpu_systems(GgrsSchedule, &[GpuSystem::ParticlesUpdate, GpuSystem::ParticlesDespawn]);
//! ```
//!
//! `handle_requests` (bevy_ggrs src/schedule_systems.rs:170-289) is replaced by [`handle_requests_b200`].
use std::ffi::{CStr, CString};

use bevy::prelude::*;
use bevy_ggrs_b200_sys as sys;
use ggrs::{Config, GgrsRequest};

/// The engine handle, a non-send resource (one caller thread, like the exclusive system that owns the World).
pub struct B200Engine(pub *mut sys::bgr_engine);

impl Drop for B200Engine {
    fn drop(&mut self) {
        unsafe { sys::bgr_engine_destroy(self.0) }
    }
}

/// Turn a non-zero status into the panic the reference would have raised, with the same text
/// (e.g. "Could not rollback to 99: no snapshot at that moment could be found.", mod.rs:209-212).
fn check(status: i32) {
    if status != sys::BGR_OK {
        let text = unsafe { CStr::from_ptr(sys::bgr_last_error()) }.to_string_lossy().into_owned();
        panic!("{text}");
    }
}

pub struct B200RollbackPlugin {
    pub max_entities: u32,
    pub max_depth: u32,
}

impl Plugin for B200RollbackPlugin {
    fn build(&self, app: &mut App) {
        let cfg = sys::bgr_config {
            abi_version: sys::BGR_ABI_VERSION,
            device: 0,
            max_entities: self.max_entities,
            max_depth: self.max_depth,
            fps: 60,
            flags: 0,
            order_base: 0,
            stream: core::ptr::null_mut(),
        };
        let mut engine = core::ptr::null_mut();
        check(unsafe { sys::bgr_engine_create(&cfg, &mut engine) });
        app.insert_non_send_resource(B200Engine(engine)).init_resource::<B200Columns>();
    }
}

/// type -> engine column id
#[derive(Resource, Default)]
pub struct B200Columns(pub bevy::platform::collections::HashMap<std::any::TypeId, u32>);

/// Registration with the reference's method names (rollback_app.rs:31-133) for POD components.
pub trait B200RollbackApp {
    fn rollback_component_with_copy_b200<T: Component + Copy + bytemuck::Pod>(&mut self) -> &mut Self;
    fn rollback_component_with_clone_b200<T: Component + Clone + bytemuck::Pod>(&mut self) -> &mut Self;
    /// Same, for a component that single entities may lose / regain inside the rollback window
    /// (`Option<&mut S::Target>` in `ComponentSnapshotPlugin::load`, component_snapshot.rs:99-115).
    fn rollback_optional_component_with_copy_b200<T: Component + Copy + bytemuck::Pod>(&mut self) -> &mut Self;
    /// `checksum_component::<T>(hasher)` where the hasher is "seahash of bytes [offset, offset+len) of T"
    fn checksum_component_b200<T: Component>(&mut self, offset: u32, len: u32, assert_finite: bool) -> &mut Self;
}

fn register<T: Component + bytemuck::Pod>(app: &mut App, strategy: u32) {
    let name = CString::new(std::any::type_name::<T>()).unwrap();
    let mut col = 0u32;
    let e = app.world().non_send_resource::<B200Engine>().0;
    check(unsafe { sys::bgr_rollback_component(e, name.as_ptr(), core::mem::size_of::<T>() as u32, strategy, &mut col) });
    app.world_mut().resource_mut::<B200Columns>().0.insert(std::any::TypeId::of::<T>(), col);
}

impl B200RollbackApp for App {
    fn rollback_component_with_copy_b200<T: Component + Copy + bytemuck::Pod>(&mut self) -> &mut Self {
        register::<T>(self, sys::BGR_STRATEGY_COPY);
        self
    }
    fn rollback_component_with_clone_b200<T: Component + Clone + bytemuck::Pod>(&mut self) -> &mut Self {
        register::<T>(self, sys::BGR_STRATEGY_CLONE);
        self
    }
    fn rollback_optional_component_with_copy_b200<T: Component + Copy + bytemuck::Pod>(&mut self) -> &mut Self {
        register::<T>(self, sys::BGR_STRATEGY_COPY | sys::BGR_STRATEGY_OPTIONAL);
        self
    }
    fn checksum_component_b200<T: Component>(&mut self, offset: u32, len: u32, assert_finite: bool) -> &mut Self {
        let col = self.world().resource::<B200Columns>().0[&std::any::TypeId::of::<T>()];
        let e = self.world().non_send_resource::<B200Engine>().0;
        let flags = if assert_finite { sys::BGR_HASH_FLAG_ASSERT_FINITE_F32 } else { 0 };
        check(unsafe { sys::bgr_checksum_component(e, col, sys::BGR_HASH_BYTES, offset, len, flags) });
        self
    }
}

/// Replacement body of `handle_requests` (schedule_systems.rs:170-289): the whole `Vec<GgrsRequest>` in ONE call.
pub fn handle_requests_b200<T: Config<Input = u8>>(requests: Vec<GgrsRequest<T>>, info: sys::bgr_session_info, world: &mut World) {
    let _span = bevy::log::tracing::info_span!("ggrs", name = "HandleRequests").entered();
    let mut cells = Vec::new();
    let reqs: Vec<sys::bgr_request> = requests
        .into_iter()
        .map(|r| match r {
            GgrsRequest::SaveGameState { cell, frame } => {
                cells.push(cell);
                sys::bgr_request { kind: sys::BGR_REQ_SAVE, frame, ..Default::default() }
            }
            GgrsRequest::LoadGameState { frame, .. } => sys::bgr_request { kind: sys::BGR_REQ_LOAD, frame, ..Default::default() },
            GgrsRequest::AdvanceFrame { inputs } => {
                let mut q = sys::bgr_request { kind: sys::BGR_REQ_ADVANCE, n_players: inputs.len() as u32, ..Default::default() };
                for (i, (input, status)) in inputs.iter().enumerate().take(sys::BGR_MAX_PLAYERS) {
                    q.inputs[i] = *input;
                    q.status[i] = *status as u8;
                }
                q
            }
        })
        .collect();
    let engine = world.non_send_resource::<B200Engine>().0;
    let mut out = [sys::bgr_checksum::default(); sys::BGR_MAX_REQUESTS];
    let mut n = 0u32;
    check(unsafe {
        sys::bgr_handle_requests(engine, &info, reqs.as_ptr(), reqs.len() as u32, out.as_mut_ptr(), out.len() as u32, &mut n)
    });
    // cell.save(frame, None, checksum)  (schedule_systems.rs:231-236) — GGRS never receives state bytes
    for (cell, cs) in cells.into_iter().zip(&out[..n as usize]) {
        cell.save(cs.frame, None, Some(((cs.hi as u128) << 64) | cs.lo as u128));
    }
    // mirror the frame resources back for user systems that read them
    let (mut frame, mut confirmed) = (0i32, 0i32);
    unsafe {
        sys::bgr_rollback_frame_count(engine, &mut frame);
        sys::bgr_confirmed_frame_count(engine, &mut confirmed);
    }
    world.insert_resource(bevy_ggrs::RollbackFrameCount(frame));
    world.insert_resource(bevy_ggrs::ConfirmedFrameCount(confirmed));
}

/// ECS table column -> HBM planes, once after spawning (Startup) or whenever the host edits a component.
/// `stride` = `size_of::<T>()` on the Rust side; the engine transposes into its tile-planar image on the GPU.
pub fn upload_column<T: Component + bytemuck::Pod>(world: &mut World, first_row: u32, values: &[T]) {
    let col = world.resource::<B200Columns>().0[&std::any::TypeId::of::<T>()];
    let engine = world.non_send_resource::<B200Engine>().0;
    check(unsafe {
        sys::bgr_write_component(engine, col, first_row, values.len() as u32, values.as_ptr().cast(), core::mem::size_of::<T>() as u32)
    });
}

/// HBM planes -> a host slice (e.g. `Transform.translation` for rendering in PostUpdate).
pub fn download_column<T: Component + bytemuck::Pod>(world: &World, first_row: u32, out: &mut [T]) {
    let col = world.resource::<B200Columns>().0[&std::any::TypeId::of::<T>()];
    let engine = world.non_send_resource::<B200Engine>().0;
    check(unsafe {
        sys::bgr_read_component(engine, col, first_row, out.len() as u32, out.as_mut_ptr().cast(), core::mem::size_of::<T>() as u32)
    });
}

/// `commands.entity(e).remove::<T>()` for a component registered with `rollback_optional_component_with_copy_b200`
/// (`row` = the entity's `RollbackOrdered` index).
pub fn remove_component<T: Component>(world: &World, row: u32) {
    let col = world.resource::<B200Columns>().0[&std::any::TypeId::of::<T>()];
    check(unsafe { sys::bgr_remove_component(world.non_send_resource::<B200Engine>().0, col, row) });
}

/// `commands.entity(e).insert(value)` for an optional component.
pub fn insert_component<T: Component + bytemuck::Pod>(world: &World, row: u32, value: &T) {
    let col = world.resource::<B200Columns>().0[&std::any::TypeId::of::<T>()];
    check(unsafe { sys::bgr_insert_component(world.non_send_resource::<B200Engine>().0, col, row, (value as *const T).cast()) });
}

/// Page-locked double buffer for the per-tick mirror of one field range of a component (INTEGRATION.md "Per-tick mirror").
pub struct B200Mirror {
    pub bufs: [*mut u8; 2],
    pub bytes_per_row: u32,
    pub rows: u32,
    pub in_flight: Option<(u32, usize)>, // (ticket, buffer index)
}

impl B200Mirror {
    pub fn new(rows: u32, bytes_per_row: u32) -> Self {
        let mut bufs = [core::ptr::null_mut::<u8>(); 2];
        for b in bufs.iter_mut() {
            let mut p: *mut core::ffi::c_void = core::ptr::null_mut();
            check(unsafe { sys::bgr_host_alloc(rows as usize * bytes_per_row as usize, &mut p) });
            *b = p.cast();
        }
        Self { bufs, bytes_per_row, rows, in_flight: None }
    }
    /// End of `run_ggrs_schedules`: start mirroring bytes [offset, offset + bytes_per_row) of every `T`; returns at once.
    pub fn begin<T: Component>(&mut self, world: &World, offset: u32, frame: usize) {
        let col = world.resource::<B200Columns>().0[&std::any::TypeId::of::<T>()];
        let mut ticket = 0u32;
        let idx = frame & 1;
        check(unsafe {
            sys::bgr_download_begin(world.non_send_resource::<B200Engine>().0, col, offset, self.bytes_per_row, 0, self.rows,
                                    self.bufs[idx].cast(), &mut ticket)
        });
        self.in_flight = Some((ticket, idx));
    }
    /// Before the first host-side reader (e.g. sprite extraction): the previous `begin`'s bytes are now readable.
    pub fn wait(&mut self, world: &World) -> Option<&[u8]> {
        let (ticket, idx) = self.in_flight.take()?;
        check(unsafe { sys::bgr_download_wait(world.non_send_resource::<B200Engine>().0, ticket) });
        Some(unsafe { core::slice::from_raw_parts(self.bufs[idx], self.rows as usize * self.bytes_per_row as usize) })
    }
}

impl Drop for B200Mirror {
    fn drop(&mut self) {
        for b in self.bufs {
            unsafe { sys::bgr_host_free(b.cast()) };
        }
    }
}

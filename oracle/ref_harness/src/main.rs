//! Reference harness: the real bevy_ggrs (+ ggrs, seahash, rand, bevy_time) driving the stress-test world headless.
//!
//! TEST INFRASTRUCTURE (oracle side).  It exists to pin the CPU restatement in `oracle/` — and through it the CUDA
//! engine — to the reference itself instead of to our reading of it:
//!
//!   * the world is the particles example's rollback registration (examples/stress_tests/particles.rs:186-222):
//!     Transform (clone), Velocity (copy, `Hash` over `to_bits`), Ttl (copy), ParticleRng (resource, clone),
//!     `checksum_component_with_hash::<Velocity>()`, `checksum_component::<Transform>(translation bits)`,
//!     GgrsSchedule = (spawn_particles.run_if(spawn_pressed), update_particles, despawn_particles);
//!   * the app is the reference's own headless test rig (tests/common/mod.rs:44-54): MinimalPlugins, a manual
//!     1/fps time step, a SyncTest session, GgrsPlugin, a ReadInputs system;
//!   * the population comes from a file written by `tests/golden/gen_reference_inputs.py` (the same seeded generator
//!     the GPU / oracle tests use), spawned in file order so that RollbackOrdered index == row.
//!
//! Output (one JSON document on stdout):
//!   checksums   [[frame, "0x<u128>"], ...]      every `Checksum` the reference hands to `cell.save`, in order
//!   dt_bits     [u32, ...]                      Time<GgrsTime>::delta_secs().to_bits() of every AdvanceWorld run
//!   final       {alive: [u8], translation: [[u32;3]], velocity: [[u32;3]], ttl: [u64]}  by RollbackOrdered index
//!   rng         first 16 `random_range(-200.0..200.0)` bit patterns of Xoshiro256PlusPlus::seed_from_u64(123)
//!   timing      {ticks, seconds, rollback_frames_per_s}   (the real reference's CPU SyncTest path: the baseline
//!               BASELINE.md §2(3) promised; bench.py --impl reference uses it when cargo is present)
//!
//! usage: bevy_ggrs_ref_harness <particles.bin> <entities> <check_distance> <ticks> [fps=60] [spawn_rate=0]

use bevy::{platform::collections::HashMap, prelude::*, time::TimeUpdateStrategy};
use bevy_ggrs::{prelude::*, checksum_hasher, Checksum, LocalInputs, LocalPlayers, RollbackFrameCount, SaveWorld, SaveWorldSystems,
                AdvanceWorld, AdvanceWorldSystems, GgrsTime, RollbackOrdered, RollbackId};
use core::time::Duration;
use ggrs::{PlayerType, SessionBuilder};
use rand::{Rng, SeedableRng};
use std::hash::{Hash, Hasher};
use std::io::Read;

type Cfg = GgrsConfig<u8>;
const INPUT_SPAWN: u8 = 1 << 4;
const INPUT_NOOP: u8 = 1 << 5;

#[derive(Component, Clone, Copy, Default)]
struct Velocity(Vec3);
impl Hash for Velocity {
    fn hash<H: Hasher>(&self, state: &mut H) {
        assert!(self.0.is_finite(), "Hashing is not stable for NaN f32 values.");
        self.0.x.to_bits().hash(state);
        self.0.y.to_bits().hash(state);
        self.0.z.to_bits().hash(state);
    }
}
#[derive(Component, Clone, Copy, Default)]
struct Ttl(usize);
#[derive(Resource, Clone)]
struct ParticleRng(rand_xoshiro::Xoshiro256PlusPlus);

#[derive(Resource)]
struct Harness { fps: usize, rate: u32, tick: u32, spawn_inputs: bool }
#[derive(Resource, Default)]
struct Log { checksums: Vec<(i32, u128)>, dt_bits: Vec<u32> }
#[derive(Resource)]
struct Population(Vec<([f32; 10], [f32; 3], u64)>);

fn read_inputs(mut commands: Commands, players: Res<LocalPlayers>, mut h: ResMut<Harness>) {
    let mut inputs = HashMap::new();
    for &handle in &players.0 {
        let mut v: u8 = if (h.tick as usize + handle) % 3 == 0 { INPUT_NOOP } else { 0 };
        if h.spawn_inputs && handle == 0 && matches!(h.tick % 5, 1 | 2) { v |= INPUT_SPAWN; }
        inputs.insert(handle, v);
    }
    h.tick += 1;
    commands.insert_resource(LocalInputs::<Cfg>(inputs));
}

fn populate(mut commands: Commands, pop: Res<Population>) {
    for (tf, vel, ttl) in pop.0.iter() {
        commands.spawn((
            Transform {
                translation: Vec3::new(tf[0], tf[1], tf[2]),
                rotation: Quat::from_xyzw(tf[3], tf[4], tf[5], tf[6]),
                scale: Vec3::new(tf[7], tf[8], tf[9]),
            },
            Velocity(Vec3::new(vel[0], vel[1], vel[2])),
            Ttl(*ttl as usize),
            Rollback,
        ));
    }
}

fn spawn_pressed(inputs: Res<PlayerInputs<Cfg>>) -> bool { inputs.iter().any(|(i, _)| *i & INPUT_SPAWN != 0) }
fn spawn_particles(mut commands: Commands, h: Res<Harness>, mut rng: ResMut<ParticleRng>) {
    let s = 200.0;
    let ttl = h.fps * 5;
    for _ in 0..h.rate {
        commands.spawn((Transform::default(), Velocity(Vec3::new(rng.0.random_range(-s..s), rng.0.random_range(-s..s), 0.0)), Ttl(ttl), Rollback));
    }
}
fn update_particles(mut q: Query<(&mut Transform, &mut Velocity)>, time: Res<Time>) {
    let dt = time.delta_secs();
    let gravity = Vec3::NEG_Y * 200.0;
    for (mut t, mut v) in &mut q {
        v.0 += gravity * dt;
        t.translation += v.0 * dt;
    }
}
fn despawn_particles(mut commands: Commands, mut q: Query<(Entity, &mut Ttl)>) {
    for (e, mut ttl) in &mut q {
        ttl.0 -= 1;
        if ttl.0 == 0 { commands.entity(e).despawn(); }
    }
}

// after ChecksumPlugin::update, inside SaveWorld: the value handle_requests passes to cell.save (schedule_systems.rs:231-236)
fn log_checksum(frame: Res<RollbackFrameCount>, checksum: Res<Checksum>, mut log: ResMut<Log>) { log.checksums.push((frame.0, checksum.0)); }
// after GgrsTimePlugin::update, inside AdvanceWorld
fn log_dt(time: Res<Time<GgrsTime>>, mut log: ResMut<Log>) { log.dt_bits.push(time.delta_secs().to_bits()); }

fn main() {
    let a: Vec<String> = std::env::args().collect();
    if a.len() < 5 { eprintln!("usage: {} <particles.bin> <entities> <check_distance> <ticks> [fps] [spawn_rate]", a[0]); std::process::exit(2); }
    let (n, d, ticks): (usize, usize, u32) = (a[2].parse().unwrap(), a[3].parse().unwrap(), a[4].parse().unwrap());
    let fps: usize = a.get(5).map(|s| s.parse().unwrap()).unwrap_or(60);
    let rate: u32 = a.get(6).map(|s| s.parse().unwrap()).unwrap_or(0);
    let mut bytes = Vec::new();
    std::fs::File::open(&a[1]).unwrap().read_to_end(&mut bytes).unwrap();
    assert_eq!(bytes.len(), n * 60, "particles.bin = n x (10 f32 transform | 3 f32 velocity | u64 ttl)");
    let f = |o: usize| f32::from_le_bytes(bytes[o..o + 4].try_into().unwrap());
    let pop = (0..n).map(|i| {
        let b = i * 60;
        let mut tf = [0f32; 10]; for k in 0..10 { tf[k] = f(b + 4 * k); }
        let vel = [f(b + 40), f(b + 44), f(b + 48)];
        (tf, vel, u64::from_le_bytes(bytes[b + 52..b + 60].try_into().unwrap()))
    }).collect();

    let session = SessionBuilder::<Cfg>::new()
        .with_num_players(2).unwrap()
        .with_check_distance(d)
        .with_max_prediction_window(d + 1)
        .with_input_delay(2)
        .add_player(PlayerType::Local, 0).unwrap()
        .add_player(PlayerType::Local, 1).unwrap()
        .start_synctest_session().unwrap();

    let mut app = App::new();
    app.add_plugins(MinimalPlugins)
        .insert_resource(TimeUpdateStrategy::ManualDuration(Duration::from_secs_f64(1.0 / fps as f64)))
        .insert_resource(RollbackFrameRate(fps))
        .add_plugins(GgrsPlugin::<Cfg>::default())
        .add_systems(ReadInputs, read_inputs)
        .rollback_component_with_clone::<Transform>()
        .rollback_component_with_copy::<Velocity>()
        .rollback_component_with_copy::<Ttl>()
        .rollback_resource_with_clone::<ParticleRng>()
        .checksum_component_with_hash::<Velocity>()
        .checksum_component::<Transform>(|t| {
            let mut hasher = checksum_hasher();
            assert!(t.translation.is_finite(), "Hashing is not stable for NaN f32 values.");
            t.translation.x.to_bits().hash(&mut hasher);
            t.translation.y.to_bits().hash(&mut hasher);
            t.translation.z.to_bits().hash(&mut hasher);
            hasher.finish()
        })
        .insert_resource(Harness { fps, rate, tick: 0, spawn_inputs: rate > 0 })
        .insert_resource(Population(pop))
        .insert_resource(ParticleRng(rand_xoshiro::Xoshiro256PlusPlus::seed_from_u64(123)))
        .init_resource::<Log>()
        .add_systems(Startup, populate)
        .add_systems(SaveWorld, log_checksum.in_set(SaveWorldSystems::Snapshot))
        .add_systems(AdvanceWorld, log_dt.in_set(AdvanceWorldSystems::Main))
        .add_observer(|ev: On<SyncTestMismatch>| panic!("SyncTestMismatch in the reference itself: {:?}", ev.event().mismatched_frames))
        .insert_resource(Session::SyncTest(session));
    if rate > 0 {
        app.add_systems(GgrsSchedule, (spawn_particles.run_if(spawn_pressed), update_particles, despawn_particles));
    } else {
        app.add_systems(GgrsSchedule, (update_particles, despawn_particles));
    }

    app.update();                                  // Startup + the zero-delta first update (no GGRS tick)
    let t0 = std::time::Instant::now();
    for _ in 0..ticks { app.update(); }
    let secs = t0.elapsed().as_secs_f64();

    // ---- dump ----
    let world = app.world_mut();
    let log = world.remove_resource::<Log>().unwrap();
    let n_adv = log.dt_bits.len();
    let ordered = world.resource::<RollbackOrdered>().clone();
    let total = ordered.len();
    let mut alive = vec![0u8; total];
    let mut tr = vec![[0u32; 3]; total];
    let mut ve = vec![[0u32; 3]; total];
    let mut tt = vec![0u64; total];
    let mut q = world.query::<(&RollbackId, &Transform, &Velocity, &Ttl)>();
    for (id, t, v, l) in q.iter(world) {
        let i = ordered.order(*id) as usize;
        alive[i] = 1;
        tr[i] = [t.translation.x.to_bits(), t.translation.y.to_bits(), t.translation.z.to_bits()];
        ve[i] = [v.0.x.to_bits(), v.0.y.to_bits(), v.0.z.to_bits()];
        tt[i] = l.0 as u64;
    }
    let mut rng = rand_xoshiro::Xoshiro256PlusPlus::seed_from_u64(123);
    let rng_bits: Vec<u32> = (0..16).map(|_| rng.random_range(-200.0f32..200.0f32).to_bits()).collect();
    let cs: Vec<String> = log.checksums.iter().map(|(f, c)| format!("[{},\"{:#x}\"]", f, c)).collect();
    println!("{{\"entities\":{},\"check_distance\":{},\"ticks\":{},\"fps\":{},\"spawn_rate\":{},", n, d, ticks, fps, rate);
    println!("\"checksums\":[{}],", cs.join(","));
    println!("\"dt_bits\":{:?},", log.dt_bits);
    println!("\"rng\":{:?},", rng_bits);
    println!("\"final\":{{\"alive\":{:?},\"translation\":{:?},\"velocity\":{:?},\"ttl\":{:?}}},", alive, tr, ve, tt);
    println!("\"timing\":{{\"ticks\":{},\"seconds\":{},\"advance_frames\":{},\"rollback_frames_per_s\":{}}}}}", ticks, secs, n_adv, n_adv as f64 / secs);
}

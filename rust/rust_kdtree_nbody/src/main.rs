//! The reference's CLI (Parallel/RustVersion/src/main.rs:8-32): `--number/-n N` (required), `--steps/-s S`
//! (default 1), dt = 1e-3, prints the elapsed seconds of initial conditions + simulation.
use std::time::Instant;

use rust_kdtree_nbody::{array_kd_tree, array_particle};

struct Args {
    number: usize,
    steps: i64,
}

fn usage() -> ! {
    eprintln!("usage: rust_kdtree_nbody --number <N> [--steps <S>]   (short forms: -n, -s)");
    std::process::exit(2)
}

fn parse_args() -> Args {
    let mut number: Option<usize> = None;
    let mut steps: i64 = 1;
    let mut it = std::env::args().skip(1);
    while let Some(arg) = it.next() {
        // both `--number 10` and `--number=10`
        let (key, inline) = match arg.split_once('=') {
            Some((k, v)) => (k.to_string(), Some(v.to_string())),
            None => (arg.clone(), None),
        };
        let mut value = || inline.clone().or_else(|| it.next()).unwrap_or_else(|| usage());
        match key.as_str() {
            "-n" | "--number" => number = Some(value().parse().unwrap_or_else(|_| usage())),
            "-s" | "--steps" => steps = value().parse().unwrap_or_else(|_| usage()),
            _ => usage(),
        }
    }
    Args { number: number.unwrap_or_else(|| usage()), steps }
}

fn main() {
    let args = parse_args();
    let dt = 1e-3;
    let start = Instant::now(); // the reference starts its timer before the initial conditions too (main.rs:25)
    array_kd_tree::simple_sim(&mut array_particle::circular_orbits(args.number), dt, args.steps);
    println!("{}", start.elapsed().as_nanos() as f64 / 1e9);
}

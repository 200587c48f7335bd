//! Reference `src/sketcharg.rs:13-138`: the parameter block every `SeqSketcherT` carries and its JSON dump
//! (`sketchparams_dump.json`, same keys and enum spellings so that dumps of either crate reload in the other).
use serde::{Deserialize, Serialize};
use std::fs::File;
use std::io::{BufReader, BufWriter};
use std::path::Path;

#[derive(Copy, Clone, Serialize, Deserialize, Debug, PartialEq, Eq)]
pub enum DataType { DNA, AA }

#[derive(Copy, Clone, Serialize, Deserialize, Debug, PartialEq, Eq)]
pub enum SketchAlgo { PROB3A, SUPER, SUPER2, OPTDENS, REVOPTDENS, HLL }

#[derive(Copy, Clone, Serialize, Deserialize, Debug)]
pub struct SeqSketcherParams {
    kmer_size: usize,
    sketch_size: usize,
    algo: SketchAlgo,
    data_t: DataType,
}

impl SeqSketcherParams {
    pub fn new(kmer_size: usize, sketch_size: usize, algo: SketchAlgo, data_t: DataType) -> Self {
        SeqSketcherParams { kmer_size, sketch_size, algo, data_t }
    }
    pub fn get_kmer_size(&self) -> usize { self.kmer_size }
    pub fn get_sketch_size(&self) -> usize { self.sketch_size }
    pub fn get_algo(&self) -> SketchAlgo { self.algo }
    pub fn get_data_t(&self) -> DataType { self.data_t }

    pub fn dump_json(&self, filename: &String) -> Result<(), String> {
        let file = File::create(filename).map_err(|_| "SeqSketcher dump failed".to_string())?;
        serde_json::to_writer(BufWriter::new(file), self).map_err(|e| e.to_string())
    }
    pub fn reload_json(dirpath: &Path) -> Result<SeqSketcherParams, String> {
        let file = File::open(dirpath.join("sketchparams_dump.json")).map_err(|_| "Sketcher reload_json could not open file".to_string())?;
        serde_json::from_reader(BufReader::new(file)).map_err(|e| e.to_string())
    }
}

//! fastq.rs of the reference: only the part of its public surface that the facade still needs.  Readers
//! and writers live inside the library's file -> file runtime (`bk_run`); a caller that drives batches
//! itself uses `BatchExtractor` (parse.rs) on buffers it read any way it likes.
pub enum CompressionType {
    Bgzf,
    Gzip,
    Mgzip,
    Lz4,
    No,
}

impl CompressionType {
    /// Same table as fastq.rs:71-79, including its `--bgz` -> Mgzip / `--mgz` -> Bgzf swap.
    pub fn select(gz: &bool, bgz: &bool, mgz: &bool, lz4: &bool) -> Self {
        match (gz, bgz, mgz, lz4) {
            (true, false, false, false) => Self::Gzip,
            (false, true, false, false) => Self::Mgzip,
            (false, false, true, false) => Self::Bgzf,
            (false, false, false, true) => Self::Lz4,
            _ => CompressionType::No,
        }
    }

    /// `bk_run_args.compression`: 0 none, 1 gzip, 2 mgzip, 3 bgzf, 4 lz4
    pub(crate) fn code(&self) -> i32 {
        match self {
            Self::No => 0,
            Self::Gzip => 1,
            Self::Mgzip => 2,
            Self::Bgzf => 3,
            Self::Lz4 => 4,
        }
    }
}

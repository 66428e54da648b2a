#!/usr/bin/env perl
# perl/bin/ppp_cross_score_b200.pl -- command-line drop-in for ProPhylo's bin/ppp_cross_score.pl on the B200 path.
#
# Same options and stdout as the reference (/root/reference/bin/ppp_cross_score.pl:210-297): for two search results
# (BLAST -m 8/9 tables or hmmsearch 3 reports, any combination) every depth prefix of one is scored against the hit list
# of the other, in both directions, and the best depth of each direction is printed.  The reference writes a temporary
# profile file per depth and runs the Perl sweep on it (:471-503, :616-650, :681-699); here all depths of a direction are
# ONE GPU call through PhyloProf::B200 (XS -> libppp_b200.so).  Parsing, gi -> taxon lookups (SeqToolBox::Taxonomy, the
# reference's own module and database) and the %.3f ranking stay on the host.  There is no CPU fallback.
#
#   perl ppp_cross_score_b200.pl --file1 A --file2 B --method1 blast|hmmer --method2 blast|hmmer --dist-file taxdis.txt
#        [--start-depth N] [--end-depth N] [--skip N] [--scale F] [--trusted_cutoff S] [-l rank] [--device N | --devices 0-7]
# -w/--workspace, --threads, -k/--keep and --prob are accepted and ignored, as far as the reference ignores them too
# (it never reads --prob; there are no temporary files to keep here).
use strict;
use warnings;
use Getopt::Long;
use FindBin;
use lib "$FindBin::Bin/../lib";
use PhyloProf::B200;
use PhyloProf::B200::Host qw(parse_hmmsearch3 above_cutoff m8_rows blast_target_list hmmer_target_list blast_prefix_queries
    hmmer_prefix_queries score_prefixes best_depth);

my ( $help, $file1, $file2, $method1, $method2, $workspace, $level, $probability, $trusted_cutoff, $threads ) = ( 0, '', '', '', '', '', '', 0, 0, 1 );
my ( $start, $end, $skip, $keep, $dist_file, $scale, $device, $devices ) = ( 1, undef, 1, 0, '', 1, 0, '' );
GetOptions(
    'help|h' => \$help, 'file1|f1=s' => \$file1, 'file2|f2=s' => \$file2, 'method1|m1=s' => \$method1, 'method2|m2=s' => \$method2,
    'workspace|w=s' => \$workspace, 'level|l=s' => \$level, 'prob=s' => \$probability, 'trusted_cutoff|t=s' => \$trusted_cutoff,
    'threads=i' => \$threads, 'start-depth|s=i' => \$start, 'end-depth|e=i' => \$end, 'skip|j=i' => \$skip, 'keep|k' => \$keep,
    'dist-file|d=s' => \$dist_file, 'scale=s' => \$scale, 'device=i' => \$device, 'devices=s' => \$devices,
) or die "usage: ppp_cross_score_b200.pl --file1 A --file2 B --method1 blast|hmmer --method2 blast|hmmer --dist-file FILE\n";
die "usage: ppp_cross_score_b200.pl --file1 A --file2 B --method1 blast|hmmer --method2 blast|hmmer --dist-file FILE\n"
    if $help || !( $file1 && $file2 && $method1 && $method2 && $dist_file );
print STDERR "Method can be only BLAST or HMMER\n" if $method1 !~ /^(blast|hmmer)$/i || $method2 !~ /^(blast|hmmer)$/i;

eval { require SeqToolBox::Taxonomy; 1 } or die "ppp_cross_score_b200.pl: needs SeqToolBox::Taxonomy (the reference's lib/) on \@INC: $@";
my $taxonomy = SeqToolBox::Taxonomy->new();
my $collapse;
if ($level) {    # ppp.pm:328-370 / Profile.pm:516-549: the rank, else genus, else the raw id
    my %memo;
    $collapse = sub {
        my $t = shift;
        return $t unless $t;
        $memo{$t} = $taxonomy->collapse_taxon( $t, $level ) || $taxonomy->collapse_taxon( $t, "genus" ) || $t unless exists $memo{$t};
        return $memo{$t};
    };
}

# number of taxa = non-comment lines of the distribution file (:299-311)
open( my $df, "<", $dist_file ) or die "Can't open $dist_file\n";
my $n_taxa = grep { !/^\#/ } <$df>;
close($df);
die "Could not determine number of taxa" unless $n_taxa;

sub load {    # one search result: its target list (what the OTHER file's prefixes are scored against) and its own prefixes
    my ( $file, $method ) = @_;
    die "Can't locate $file\n" unless -s $file;
    my %o = ( start => $start, end => $end, skip => $skip, scale => $scale );
    if ( $method =~ /hmmer/i ) {
        my $hits = above_cutoff( parse_hmmsearch3($file)->{domain}, $trusted_cutoff );    # :326-348, :549-561: best-domain scores
        return { method => 'hmmer', target => hmmer_target_list( $hits, $taxonomy ), prefixes => sub { hmmer_prefix_queries( $hits, $taxonomy, $n_taxa, %o ) } };
    } elsif ( $method =~ /blast/i ) {
        my $rows = m8_rows($file);
        return { method => 'blast', target => blast_target_list( $rows, $taxonomy ), prefixes => sub { blast_prefix_queries( $rows, $taxonomy, $n_taxa, %o ) } };
    }
    die "Could not understand format option $method\n";
}

my $r1  = load( $file1, $method1 );
my $r2  = load( $file2, $method2 );
my $gpu = $devices ne '' ? PhyloProf::B200->new( -devices => $devices ) : PhyloProf::B200->new( -device => $device );
my $best_gi;    # the reference's $best_score_gi: reset by a BLAST evaluation, left alone by a HMMER one (:429, :549-679)
my @blocks;
foreach my $dir ( [ "File1", $file1, $r1, $file2, $r2, "#Total" ], [ "File2", $file2, $r2, $file1, $r1, "#Toal" ] ) {
    my ( $which, $fa, $ra, $fb, $rb, $total ) = @$dir;
    my $queries = $ra->{prefixes}->();
    my $best = best_depth( $queries, score_prefixes( $gpu, $queries, $rb->{target}, $collapse ) );
    die "no depth of $fa was evaluated\n" unless $best;
    $best_gi = $best->[1] if $ra->{method} eq 'blast';
    push @blocks, "$which: $fa Best result in $fb:\nSource_Line:gi\tScore\t#yes\t$total\t#Depth\t#Prob\n" . $best->[0] . ":" . ( defined $best_gi ? $best_gi : "" ) . "\t"
        . join( "\t", @{ $best->[2] } ) . "\n";
}
print $blocks[0], "\n", $blocks[1];
exit(0);
